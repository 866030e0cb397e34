#!/usr/bin/env python
"""Static issue-time estimate of a SASS region from the control codes (dev tool, no GPU needed).

  cuobjdump -sass -fun <mangled> build/x.o | python tools/sass_stalls.py [--from ADDR --to ADDR]

Every sm_100 instruction carries a stall count (cycles before the NEXT instruction of the warp may issue); their sum over a
straight-line region is the time ONE warp needs alone when every variable-latency wait (scoreboards) is already satisfied.
Comparing it with 2 x (#FP64 instructions) -- the FP64 datapath time of a warp instruction -- shows how latency-bound the
region is for a single warp."""
import re
import sys

pat = re.compile(r"/\*([0-9a-f]{4,6})\*/\s+(.*?);\s+/\* 0x([0-9a-f]{16}) \*/")
pat2 = re.compile(r"/\* 0x([0-9a-f]{16}) \*/")


def parse(lines):
    out = []
    it = iter(lines)
    for ln in it:
        m = pat.search(ln)
        if not m:
            continue
        nxt = next(it)
        m2 = pat2.search(nxt)
        hi = int(m2.group(1), 16)
        ctrl = hi >> 41
        out.append({"addr": int(m.group(1), 16), "text": m.group(2).strip(), "stall": ctrl & 0xF, "yield": (ctrl >> 4) & 1,
                    "wbar": (ctrl >> 5) & 7, "rbar": (ctrl >> 8) & 7, "wait": (ctrl >> 11) & 0x3F})
    return out


def main():
    a0 = a1 = None
    args = sys.argv[1:]
    if "--from" in args:
        a0 = int(args[args.index("--from") + 1], 16)
    if "--to" in args:
        a1 = int(args[args.index("--to") + 1], 16)
    ins = parse(sys.stdin.readlines())
    sel = [i for i in ins if (a0 is None or i["addr"] >= a0) and (a1 is None or i["addr"] < a1)]
    dp = [i for i in sel if re.match(r"(@!?U?P\d+\s+)?(DFMA|DMUL|DADD)", i["text"])]
    dmma = [i for i in sel if "DMMA" in i["text"]]
    stall = sum(max(i["stall"], 1) for i in sel)
    print(f"instructions {len(sel)}  FP64 {len(dp)}  DMMA {len(dmma)}  sum(stall) {stall}  "
          f"FP64 datapath cycles {2 * len(dp) + 16 * len(dmma)}  ratio datapath/stall {(2 * len(dp) + 16 * len(dmma)) / max(stall, 1):.2f}")
    waits = sum(1 for i in sel if i["wait"])
    print(f"instructions waiting on a scoreboard: {waits}")


if __name__ == "__main__":
    main()
