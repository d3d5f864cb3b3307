//! Drop-in replacement for `compression::bzip2::decoder::BZip2Decoder` (chalharu/rust-compression,
//! src/bzip2/decoder.rs:584-615) that keeps the crate's API (`Decoder::next`, `DecodeExt::decode`,
//! `BZip2Error`) and decodes all blocks of the buffer in parallel on a B200 through the C ABI of include/bzb200.h
//! (section 3).
//!
//! SOURCE ONLY: the build image has no rustc/cargo.  The byte sequence yielded over all `next`
//! calls is the reference decoder's; for a malformed buffer the bytes the reference yields before it fails are
//! yielded first and then the same `BZip2Error` kind.  The input iterator is drained before the first byte is
//! handed out (the reference decodes while it pulls input).

use crate::bzip2::error::BZip2Error;
use crate::traits::decoder::Decoder;
use std::os::raw::c_int;

#[repr(C)]
struct BzbDec {
    _private: [u8; 0],
}

const BZB200_OK: c_int = 0;
const BZB200_E_DATA: c_int = -6;

#[link(name = "bzb200")]
extern "C" {
    fn bzb200_dec_create(device: c_int, out: *mut *mut BzbDec) -> c_int;
    fn bzb200_dec_write(d: *mut BzbDec, p: *const u8, n: usize) -> c_int;
    fn bzb200_dec_finish(d: *mut BzbDec) -> c_int;
    fn bzb200_dec_error_kind(d: *const BzbDec) -> c_int;
    fn bzb200_dec_read(d: *mut BzbDec, dst: *mut u8, cap: usize) -> usize;
    fn bzb200_dec_reset(d: *mut BzbDec) -> c_int;
    fn bzb200_dec_destroy(d: *mut BzbDec);
}

const CHUNK: usize = 1 << 20;

#[derive(Debug)]
pub struct BZip2Decoder {
    handle: *mut BzbDec,
    outbuf: Vec<u8>,
    outpos: usize,
    decoded: bool,
}

impl BZip2Decoder {
    pub fn new() -> Self {
        let mut handle: *mut BzbDec = std::ptr::null_mut();
        let rc = unsafe { bzb200_dec_create(-1, &mut handle) };
        if rc != BZB200_OK || handle.is_null() {
            // `new` cannot fail in the reference's API (decoder.rs:584-601): without a device every `next` yields
            // Err(Unexpected)
            handle = std::ptr::null_mut();
        }
        Self { handle, outbuf: Vec::new(), outpos: 0, decoded: false }
    }

    fn kind(code: c_int) -> BZip2Error {
        match code {
            // BZB200_BZ_* of include/bzb200.h = ordinal + 1 of bzip2/error.rs:4-11
            1 => BZip2Error::DataError,
            2 => BZip2Error::DataErrorMagicFirst,
            3 => BZip2Error::DataErrorMagic,
            4 => BZip2Error::UnexpectedEof,
            _ => BZip2Error::Unexpected,
        }
    }
}

impl Default for BZip2Decoder {
    fn default() -> Self {
        Self::new()
    }
}

impl Drop for BZip2Decoder {
    fn drop(&mut self) {
        if !self.handle.is_null() {
            unsafe { bzb200_dec_destroy(self.handle) }
        }
    }
}

impl Decoder for BZip2Decoder {
    type Input = u8;
    type Output = u8;
    type Error = BZip2Error;

    // decoder.rs:607-614
    fn next<I: Iterator<Item = u8>>(&mut self, iter: &mut I) -> Option<Result<u8, BZip2Error>> {
        if self.handle.is_null() {
            return Some(Err(BZip2Error::Unexpected)); // bzb200_dec_create failed in `new`
        }
        loop {
            if self.outpos < self.outbuf.len() {
                let b = self.outbuf[self.outpos];
                self.outpos += 1;
                return Some(Ok(b));
            }
            if self.decoded {
                self.outbuf.resize(CHUNK, 0);
                let n = unsafe { bzb200_dec_read(self.handle, self.outbuf.as_mut_ptr(), CHUNK) };
                self.outbuf.truncate(n);
                self.outpos = 0;
                if n > 0 {
                    continue;
                }
                let code = unsafe { bzb200_dec_error_kind(self.handle) };
                self.decoded = false;
                unsafe { bzb200_dec_reset(self.handle) };
                return if code != 0 { Some(Err(Self::kind(code))) } else { None };
            }
            let mut inbuf: Vec<u8> = Vec::with_capacity(CHUNK);
            for b in iter.by_ref() {
                inbuf.push(b);
                if inbuf.len() == CHUNK {
                    if unsafe { bzb200_dec_write(self.handle, inbuf.as_ptr(), inbuf.len()) } != BZB200_OK {
                        return Some(Err(BZip2Error::Unexpected));
                    }
                    inbuf.clear();
                }
            }
            if !inbuf.is_empty()
                && unsafe { bzb200_dec_write(self.handle, inbuf.as_ptr(), inbuf.len()) } != BZB200_OK
            {
                return Some(Err(BZip2Error::Unexpected));
            }
            let rc = unsafe { bzb200_dec_finish(self.handle) };
            if rc != BZB200_OK && rc != BZB200_E_DATA {
                unsafe { bzb200_dec_reset(self.handle) };
                return Some(Err(BZip2Error::Unexpected)); // CUDA failure
            }
            self.decoded = true;
            self.outbuf.clear();
            self.outpos = 0;
        }
    }
}
