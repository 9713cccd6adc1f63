#!/usr/bin/env python
"""Decode the scheduling control bits of sm_100 SASS (no GPU needed): for every instruction of `cuobjdump -sass` output
print which scoreboard it sets when its result lands (wr), which it sets when its operands have been read (rd) and which
scoreboards it waits for.  ptxas has six scoreboards per warp and hands them out itself; when two unrelated long-latency
instructions share one (e.g. the z-buffer gathers and a tile-claim atomic), waiting for either waits for both — which
is what decided the deferred-gather experiment in profiles/r01j_exp_ring_dynamic.json.

    cuobjdump -sass -fun <mangled kernel> lib.so > k.sass
    tools/sass_scoreboards.py k.sass ['LDG|ATOMG|REDG|LDS|SYNCS']      (regex of the instructions to list; the ones that
                                                                         wait for anything are always listed)

Control field = bits 105..125 of the 128-bit instruction: stall[4] yield[1] wr[3] rd[3] wait[6] reuse[4] (7 = no scoreboard).
"""
import re
import sys


def decode(path):
    lines = open(path).read().split("\n")
    i = 0
    while i + 1 < len(lines):
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);\s+/\* (0x[0-9a-f]{16}) \*/", lines[i])
        m2 = re.match(r"\s+/\* (0x[0-9a-f]{16}) \*/", lines[i + 1]) if m else None
        if m2:
            ctrl = (int(m2.group(1), 16) >> 41) & 0x1FFFFF
            yield m.group(1), m.group(2).strip(), ctrl & 0xF, (ctrl >> 5) & 7, (ctrl >> 8) & 7, (ctrl >> 11) & 0x3F
            i += 2
        else:
            i += 1


def main():
    pat = sys.argv[2] if len(sys.argv) > 2 else r"LDG|ATOMG|REDG|LDS|SYNCS\.PHASE|UBLKCP"
    for addr, text, stall, wr, rd, wait in decode(sys.argv[1]):
        w = "".join(str(b) for b in range(6) if wait >> b & 1)
        if w or re.search(pat, text):
            print(f"{addr} wr={wr if wr != 7 else '-'} rd={rd if rd != 7 else '-'} wait={w or '-':6s} stall={stall:2d}  {text[:90]}")


if __name__ == "__main__":
    main()
