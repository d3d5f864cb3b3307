#!/usr/bin/env python3
"""Extract the golden vectors the reference's OWN unit tests hold for the bzip2
encode path into tests/golden/reference_vectors.json, and copy the reference's
data fixtures (data/sample1-7.ref, test inputs, not source) to tests/golden/data/.

Run in the authoring container only (needs /root/reference); the outputs are
committed because /root/reference does not exist on the GPU box.

Sources parsed (file:line of the #[test] functions):
  src/bzip2/mod.rs:41-58            test_unit      b"a\\n" @ level 9 -> 39 bytes
  src/suffix_array/sais.rs:294-556  test_bwt1..12  BWT strings + rotation-index arrays
  src/huffman/cano_huff_table.rs:238-294           code-length vectors
  src/huffman/encoder.rs:64-79      lefthuffman_encode_new  canonical codes
  src/bitio/writer.rs:253-322       leftbitwriter_*  MSB-first packing
"""
import hashlib
import json
import os
import re
import shutil

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def strip_comments(src):
    return "\n".join(l for l in src.splitlines() if not l.strip().startswith("//"))


def parse_bytes(tok):
    tok = tok.strip()
    if tok.startswith('b"'):
        body = tok[2:-1]
        assert "\\" not in body, body
        return body.encode("latin1")
    assert tok.startswith("&["), tok
    return bytes(int(x) for x in tok[2:-1].replace("\n", " ").split(",") if x.strip())


def main():
    out = {}
    sais = strip_comments(open(f"{REF}/src/suffix_array/sais.rs").read())
    sais = sais[sais.index("mod tests"):]
    arg = r'(b"[^"]*"|&\[[^\]]*\])'
    bwt_str, bwt_pos = [], []
    for m in re.finditer(r"test_bwt\(\s*" + arg + r"\s*,\s*" + arg + r"\s*,?\s*\)", sais):
        bwt_str.append({"src": parse_bytes(m.group(1)).hex(), "bwt": parse_bytes(m.group(2)).hex()})
    for m in re.finditer(r"test_bwtpos\(\s*" + arg + r"\s*,\s*&\[([^\]]*)\]\s*,?\s*\)", sais):
        pos = [int(x) for x in m.group(2).replace("\n", " ").split(",") if x.strip()]
        bwt_pos.append({"src": parse_bytes(m.group(1)).hex(), "pos": pos})
    # the helper definitions themselves match test_bwt( / test_bwtpos( only with literal args, so counts are exact
    out["bwt_str"] = bwt_str
    out["bwt_pos"] = bwt_pos

    mod = open(f"{REF}/src/bzip2/mod.rs").read()
    m = re.search(r"Ok\(vec!\[(.*?)\]\)", mod, re.S)
    unit = bytes(int(x.strip(), 16) for x in m.group(1).split(",") if x.strip())
    assert len(unit) == 39
    out["test_unit"] = {"input": b"a\n".hex(), "level": 9, "output": unit.hex()}

    # cano_huff_table.rs:252-264 / :238-250 / :266-286 / :288-294
    out["huffman"] = {
        "with_fn": {"freq": [0, 1, 1, 2, 2, 4, 4, 8, 8], "shift": 8, "lim": 12,
                    "lens": [0, 4, 4, 4, 4, 3, 3, 2, 2]},
        "cost80": {"freq": [0, 1, 1, 2, 2, 4, 4, 8, 8], "lim": 12, "total_cost": 80},
        "lim_len": {"freq": list(range(63)), "shift": 8, "lim": 8},
        "unit": {"freq": [0, 1], "lim": 12, "lens": [0, 1]},
    }
    # huffman/encoder.rs:64-79
    out["canonical"] = {"lens": [0, 4, 4, 4, 4, 3, 3, 2, 2],
                        "codes": [None, 0b1100, 0b1101, 0b1110, 0b1111, 0b100, 0b101, 0b00, 0b01]}
    # bitio/writer.rs:253-322 (Left cases)
    out["bitwriter"] = [
        {"fields": [[1, 1], [2, 2], [3, 3], [0, 2]], "bytes": [0b11001100]},
        {"fields": [[975, 10], [475, 10], [3784, 12]], "bytes": [243, 221, 190, 200]},
        {"fields": [[1, 1], [2, 2], [3, 3]], "bytes": [204]},
        {"fields": [[1, 1]], "bytes": [128]},
        {"fields": [[0, 10], [0, 0], [0, 1], [0, 2], [0, 3], [0, 4], [0, 12]], "bytes": [0, 0, 0, 0]},
    ]

    os.makedirs(f"{HERE}/data", exist_ok=True)
    files = {}
    for i in range(1, 8):
        name = f"sample{i}.ref"
        shutil.copyfile(f"{REF}/data/{name}", f"{HERE}/data/{name}")
        d = open(f"{REF}/data/{name}", "rb").read()
        files[name] = {"size": len(d), "sha256": hashlib.sha256(d).hexdigest()}
    out["data_files"] = files

    with open(f"{HERE}/reference_vectors.json", "w") as f:
        json.dump(out, f, indent=1)
    print(f"bwt_str={len(bwt_str)} bwt_pos={len(bwt_pos)} files={len(files)}")


if __name__ == "__main__":
    main()
