#!/usr/bin/env python
"""CPU-side look at the SASS of encode_kernel's cross-warp hand-offs (no GPU needed).

The defect of round 1 was ptxas hoisting a data load above the spin loop on its flag.  The
hand-offs are now acquire / release / relaxed operations of the PTX memory model
(ld.acquire.cta.shared is a plain LDS on sm_100a, so nothing in the SASS says "acquire":
what can be checked is that the loads are where the source puts them).  This prints, for
both instantiations, the shared-memory loads and branches around (1) the decision warp's
record poll -- the 128-bit load must sit INSIDE the loop, (2) the front end's tag poll --
the row/page loads must FOLLOW its exit, (3) the stream-P block wait -- the flag's load must
sit inside the loop, with its back-off.  Exits non-zero if (1) or (3) does not hold, or if
the kernel holds no MEMBAR.ALL.CTA (the releases) at all.

    python scripts/check_handoffs.py            # after python -m iivision_b200._build
"""
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "iivision_b200", "csrc", "iiv_encoder.cu")
OBJ = os.path.join(ROOT, "iivision_b200", "build", "iiv_encoder.o")


def main():
    src = open(SRC).read().split("\n")

    def line_of(pat):
        return [i + 1 for i, l in enumerate(src) if pat in l][0]
    l_poll = line_of("} while ((rec.x >> 16) != seq);")
    l_after = line_of("tag = ld_acq_u32(&tags[slot]);")
    l_mt = line_of("while ((mt_seen = (int)ld_acq_u32(&sm.mt_done)) < mt_issued - 1)")
    with tempfile.TemporaryDirectory() as tmp:
        subprocess.run(["cuobjdump", "-xelf", "all", OBJ], cwd=tmp, check=True,
                       stdout=subprocess.DEVNULL)
        cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
        dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)],
                             capture_output=True, text=True, check=True).stdout
    bad = 0
    for kern in ("encode_kernelILi0", "encode_kernelILi1"):
        active, cur, seq = False, None, []
        for ln in dis.splitlines():
            if ln.startswith("//---") and ".text." in ln:
                active = kern in ln
                continue
            if not active:
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if m:
                cur = (m.group(1).split("/")[-1], int(m.group(2)))
                continue
            m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
            if m:
                seq.append((int(m.group(1), 16), cur, m.group(2)[:80]))
            elif re.match(r"\s*\.L_x_\d+:", ln):
                seq.append((None, cur, ln.strip()))
        print("=====", kern)
        for name, line, span in (("record poll", l_poll, (14, 10)),
                                 ("front-end loads after the tag", l_after, (30, 22)),
                                 ("stream-P block wait", l_mt, (8, 14))):
            idxs = [i for i, (a, c, _) in enumerate(seq)
                    if c and c[0] == "iiv_encoder.cu" and abs(c[1] - line) <= 3 and a is not None]
            lo, hi = min(idxs) - span[0], max(idxs) + span[1]
            window = seq[lo:hi]
            print("--", name)
            for a, c, ins in window:
                if any(k in ins for k in ("LDS", "BRA", "L_x", "MEMBAR", "NANOSLEEP")):
                    print("   %s %s" % ("%05x" % a if a is not None else "     ", ins))
            text = [ins for _, _, ins in window]
            if name == "record poll":
                # label, LDS.128, ..., backward branch to that label
                labels = [i for i, t in enumerate(text) if t.startswith(".L_x_")]
                ok = any(any("LDS.128" in t for t in text[i:j]) and
                         any("BRA `(%s)" % text[i].rstrip(":") in t for t in text[i:j])
                         for i in labels for j in (min(i + 8, len(text)),))
                if not ok:
                    print("   !! no 128-bit load inside the poll loop")
                    bad += 1
            if name == "stream-P block wait":
                labels = [i for i, t in enumerate(text) if t.startswith(".L_x_")]
                ok = any(any(t.startswith("LDS") or " LDS" in t for t in text[i:j]) and
                         any("NANOSLEEP" in t for t in text[i:j]) and
                         any("BRA `(%s)" % text[i].rstrip(":") in t for t in text[i:j])
                         for i in labels for j in (min(i + 10, len(text)),))
                if not ok:
                    print("   !! the flag is not re-loaded inside the wait loop")
                    bad += 1
        n_rel = sum(1 for _, _, ins in seq if "MEMBAR.ALL.CTA" in ins)
        n_sc = sum(1 for _, _, ins in seq if "MEMBAR.SC" in ins)
        print("-- fences: %d MEMBAR.ALL.CTA (release), %d MEMBAR.SC.*" % (n_rel, n_sc))
        if n_rel == 0:
            print("   !! no release fence in the kernel")
            bad += 1
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
