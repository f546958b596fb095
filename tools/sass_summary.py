"""profiles/r2_sass_excerpts.md: per-kernel counts of the tcgen05 / TMEM / TMA / packed-fp32 SASS mnemonics
in the built library (cuobjdump -sass; runs on the CPU box)."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "latteclip_b200", "_C", "liblatte_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)
keys = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTMAPF", "SYNCS", "FFMA2", "FADD2",
        "FMUL2", "MUFU.EX2", "REDG", "MEMBAR", "ELECT", "UCGABAR", "UBLKCP", "NANOSLEEP.SYNCS"]
show = ("UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "FFMA2", "UTCBAR", "UBLKCP", "NANOSLEEP.SYNCS")
out = ["# SASS evidence, round 2", "",
       "`cuobjdump -sass latteclip_b200/_C/liblatte_b200.so` (sm_100a, the library the tests and bench load), per kernel:",
       "instruction count, counts of the tcgen05 (`UTCHMMA`, `UTCBAR`), TMEM (`LDTM`, `STTM`), TMA (`UTMALDG`, `UTMASTG`,",
       "`UTMAREDG`, 1-D bulk copies `UBLKCP`) and packed-fp32 (`FFMA2`, `FADD2`, `FMUL2`) mnemonics, the suspended mbarrier",
       "wait (`NANOSLEEP.SYNCS`, DESIGN 3.10), and the first occurrence of the main ones.",
       "Regenerate with `python tools/sass_summary.py`.", ""]
for f in funcs[1:]:
    name = f.split("\n", 1)[0].strip()
    lines = [l for l in f.split("\n") if re.search(r"/\*[0-9a-f]{4,6}\*/\s+\S", l)]
    dn = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    short = re.sub(r"latte::\(anonymous namespace\)::", "", dn).split("(")[0]
    cnt, first = collections.OrderedDict(), {}
    for l in lines:
        m = re.search(r"/\*[0-9a-f]{4,6}\*/\s+(.*?);", l)
        if not m:
            continue
        ins = m.group(1)
        for k in keys:
            if k in ins:
                cnt[k] = cnt.get(k, 0) + 1
                first.setdefault(k, ins.strip())
    out.append(f"## {short}   ({len(lines)} SASS instructions)")
    if cnt:
        out.append("    " + "  ".join(f"{k}={v}" for k, v in cnt.items()))
        out += [f"        {v}" for k, v in first.items() if k in show]
    out.append("")
open(os.path.join(ROOT, "profiles", "r2_sass_excerpts.md"), "w").write("\n".join(out))
print(len(funcs) - 1, "kernels")
