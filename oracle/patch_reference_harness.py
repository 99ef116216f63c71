#!/usr/bin/env python
"""Generates the reference's own harness with the B200 decoder rows — TEST INFRASTRUCTURE ONLY.

Reads <ref>/src/main.cpp where it lies, applies exactly the edit INTEGRATION.md §1 asks a maintainer to make, and writes
the result to the path given (under the git-ignored oracle/_ref/; reference text is never committed here):

  * `#include "hsrans_b200_codecs.hpp"` after the reference's codec headers (src/main.cpp:7-16);
  * one more `decoders[]` entry at the end of every `_Codecs[]` row (src/main.cpp:174-228) whose codec the B200
    library decodes: `{ "dec B200", cuda_<reference entry point>_<bits>, true }`;
  * `MaxDecoderCount` 32 -> 34 (src/main.cpp:143, "if someone needs more than 32, simply increase this"): the
    rANS32x32 (raw) rows for 10..12 bits already list 31 decoders plus the terminator;
  * for the mt_ rows also `{ "dec B200 (pool signature)", decode_with_thread_pool_wrapper<cuda_mt_..._decode_mt_<b>>, true }`,
    i.e. the reference's own wrapper template (src/main.cpp:163-170) instantiated on the pool-signature twin.

Nothing else changes: the reference's encoders produce the streams, its loop times the rows, its Validate() compares
every decoded byte (src/main.cpp:817-835,860-897,949-1039), `--test` exits non-zero on the first mismatch (:359-371).
The rows are appended, so decoders[0] — the one the harness validates each ENCODER with — stays the reference's.

    python oracle/patch_reference_harness.py /root/reference/src/main.cpp oracle/_ref/harness/main_b200.cpp
"""
import re
import sys

ROW = re.compile(r'^\s*\{ "(?P<name>[^"]+)", (?P<bits>1[0-5]), ')
NAMES = {
    "rANS32x32 16w (variable block size)": ("cuda_block_rANS32x32_16w_decode", None),
    "rANS32x64 16w (variable block size)": ("cuda_block_rANS32x64_16w_decode", None),
    "rANS32x32 16w (independent blocks)": ("cuda_mt_rANS32x32_16w_decode", "cuda_mt_rANS32x32_16w_decode_mt"),
    "rANS32x64 16w (independent blocks)": ("cuda_mt_rANS32x64_16w_decode", "cuda_mt_rANS32x64_16w_decode_mt"),
    "rANS32x32 16w (raw)": ("cuda_rANS32x32_16w_decode", None),
    "rANS32x64 16w (raw)": ("cuda_rANS32x64_16w_decode", None),
    "rANS32x16 16w (raw)": ("cuda_rANS32x16_16w_decode", None),
    "rANS32x32 32blk 16w (raw)": ("cuda_rANS32x32_32blk_16w_decode", None),
}
TERMINATOR = ", {}}},"  # end of the decoders[] initialiser and of the row


def patch(text: str):
    out, rows, in_table, included, bumped = [], 0, False, False, False
    for line in text.split("\n"):
        if not included and line.startswith('#include "mt_rANS32x64_16w.h"'):
            out.append(line)
            out.append('#include "hsrans_b200_codecs.hpp" // B200 rows (INTEGRATION.md §1)')
            included = True
            continue
        if line.startswith("constexpr size_t MaxDecoderCount = 32;"):
            line = line.replace("= 32;", "= 34;", 1)
            bumped = True
        if line.startswith("static const codec_info_t _Codecs[]"):
            in_table = True
        elif in_table and line.startswith("};"):
            in_table = False
        m = ROW.match(line) if in_table else None
        if m and m.group("name") in NAMES:
            fn, pool_fn = NAMES[m.group("name")]
            bits = m.group("bits")
            stripped = line.rstrip()
            if not stripped.endswith(TERMINATOR):
                raise SystemExit(f"unexpected row shape: {stripped[:80]}...")
            if stripped.split("{}}, {{", 1)[1].count('{ "') + 1 + 2 + 1 > 34:  # present + ours + terminator vs MaxDecoderCount
                raise SystemExit("row has no room for more decoders")
            extra = f', {{ "dec B200", {fn}_{bits}, true }}'
            if pool_fn:
                extra += f', {{ "dec B200 (pool signature)", decode_with_thread_pool_wrapper<{pool_fn}_{bits}>, true }}'
            line = stripped[: -len(TERMINATOR)] + extra + TERMINATOR
            rows += 1
        out.append(line)
    if not included or not bumped:
        raise SystemExit("include / MaxDecoderCount anchor not found")
    if rows != 48:
        raise SystemExit(f"expected 48 codec rows to extend, found {rows}")
    return "\n".join(out)


if __name__ == "__main__":
    src, dst = sys.argv[1], sys.argv[2]
    with open(src, encoding="utf-8", errors="surrogateescape") as f:
        patched = patch(f.read())
    with open(dst, "w", encoding="utf-8", errors="surrogateescape") as f:
        f.write(patched)
    print(f"{dst}: 48 rows extended")
