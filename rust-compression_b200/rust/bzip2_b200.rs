//! Drop-in replacement for `compression::bzip2::encoder::BZip2Encoder` (chalharu/rust-compression,
//! src/bzip2/encoder.rs:39-159) that keeps the crate's API and runs the block-compression hot path on a B200
//! through the C ABI of include/bzb200.h.
//!
//! SOURCE ONLY: the build image has no rustc/cargo, so this file is shipped uncompiled (INTEGRATION.md shows
//! where it goes in the crate).  The byte sequence produced over all `next` calls is identical to the
//! reference encoder's; closed blocks are compressed whenever a window of input has accumulated, the rest at
//! `Action::Finish`.  The decode side (`BZip2Decoder`) is in bzip2_decoder_b200.rs.
//!
//! In the crate: replace `pub use crate::bzip2::encoder::BZip2Encoder` (src/lib.rs:77) by this type behind a
//! `b200` cargo feature and link with `-lbzb200`.

use crate::action::Action;
use crate::error::CompressionError;
use crate::traits::encoder::Encoder;
use std::os::raw::{c_char, c_int};

#[repr(C)]
struct BzbEnc {
    _private: [u8; 0],
}

#[link(name = "bzb200")]
extern "C" {
    fn bzb200_enc_create(level: c_int, device: c_int, out: *mut *mut BzbEnc) -> c_int;
    fn bzb200_enc_write(e: *mut BzbEnc, p: *const u8, n: usize) -> c_int;
    fn bzb200_enc_finish(e: *mut BzbEnc) -> c_int;
    fn bzb200_enc_read(e: *mut BzbEnc, dst: *mut u8, cap: usize) -> usize;
    fn bzb200_enc_reset(e: *mut BzbEnc) -> c_int;
    fn bzb200_enc_destroy(e: *mut BzbEnc);
    #[allow(dead_code)]
    fn bzb200_enc_last_error(e: *const BzbEnc) -> *const c_char;
}

const CHUNK: usize = 1 << 20;

#[derive(Debug)]
pub struct BZip2Encoder {
    handle: *mut BzbEnc,
    inbuf: Vec<u8>,
    outbuf: Vec<u8>,
    outpos: usize,
    finished: bool, // mirrors encoder.rs:45/48 (`finished` / `bit_finished` latches)
}

impl Default for BZip2Encoder {
    fn default() -> Self {
        Self::new(9) // encoder.rs:51-55
    }
}

impl BZip2Encoder {
    pub fn new(level: usize) -> Self {
        if level < 1 || level > 9 {
            panic!("invalid level"); // encoder.rs:59-61
        }
        let mut handle: *mut BzbEnc = std::ptr::null_mut();
        let rc = unsafe { bzb200_enc_create(level as c_int, -1, &mut handle) };
        if rc != 0 || handle.is_null() {
            // no CUDA device / library error: the constructor cannot fail in the reference's API (encoder.rs:58-72),
            // so the object is built without a handle and every `next` yields Err(Unexpected), the one error the
            // reference's encoder surfaces (encoder.rs:623)
            handle = std::ptr::null_mut();
        }
        Self {
            handle,
            inbuf: Vec::with_capacity(CHUNK),
            outbuf: vec![0u8; CHUNK],
            outpos: 0,
            finished: false,
        }
    }

    fn push_input(&mut self) -> Result<(), CompressionError> {
        if !self.inbuf.is_empty() {
            let rc = unsafe { bzb200_enc_write(self.handle, self.inbuf.as_ptr(), self.inbuf.len()) };
            self.inbuf.clear();
            if rc != 0 {
                return Err(CompressionError::Unexpected);
            }
        }
        Ok(())
    }

    fn refill(&mut self) -> usize {
        let n = unsafe { bzb200_enc_read(self.handle, self.outbuf.as_mut_ptr(), CHUNK) };
        self.outbuf.truncate(0);
        unsafe { self.outbuf.set_len(n) };
        self.outpos = 0;
        n
    }
}

impl Drop for BZip2Encoder {
    fn drop(&mut self) {
        if !self.handle.is_null() {
            unsafe { bzb200_enc_destroy(self.handle) }
        }
    }
}

impl Encoder for BZip2Encoder {
    type Error = CompressionError;
    type In = u8;
    type Out = u8;

    // encoder.rs:120-158: one output byte per call; None = drained.  Under Action::Run the library compresses the
    // blocks that have closed whenever a window of input has accumulated (bzb200_enc_write), so bytes can become
    // available before Finish, as in the reference (encoder.rs:91-107).
    fn next<I: Iterator<Item = u8>>(&mut self, iter: &mut I, action: Action) -> Option<Result<u8, CompressionError>> {
        if self.handle.is_null() {
            return Some(Err(CompressionError::Unexpected)); // bzb200_enc_create failed in `new`
        }
        loop {
            if self.outpos < self.outbuf.len() {
                let b = self.outbuf[self.outpos];
                self.outpos += 1;
                return Some(Ok(b));
            }
            if !self.finished {
                // drain the caller's iterator (the reference pulls one byte at a time, :79-85)
                for b in iter.by_ref() {
                    self.inbuf.push(b);
                    if self.inbuf.len() == CHUNK {
                        if let Err(e) = self.push_input() {
                            return Some(Err(e));
                        }
                    }
                }
                if let Err(e) = self.push_input() {
                    return Some(Err(e));
                }
                if let Action::Finish = action {
                    if unsafe { bzb200_enc_finish(self.handle) } != 0 {
                        return Some(Err(CompressionError::Unexpected)); // encoder.rs:623 is the only error the reference yields
                    }
                    self.finished = true;
                }
                // Flush is out of contract and behaves like Run.
            }
            self.outbuf.resize(CHUNK, 0);
            if self.refill() > 0 {
                continue;
            }
            self.outbuf.clear();
            if self.finished {
                // stream fully handed out: re-arm like encoder.rs:87-90,130-133
                self.finished = false;
                unsafe { bzb200_enc_reset(self.handle) };
            }
            return None; // Run: "input drained, feed more" (:106,142-144)
        }
    }
}
