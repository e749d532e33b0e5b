"""Per-kernel SASS evidence (B200_PROFILING.md, "What proves a Blackwell-native kernel"): counts of the mnemonics that
tcgen05 / TMEM / TMA / legacy-MMA / cluster / async-copy code compiles to, for every kernel in libeda_b200.so.
CPU-only (cuobjdump on the built library).  Usage: python scripts/sass_evidence.py [out.json]"""
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "eda_b200", "lib", "libeda_b200.so")
PAT = {
    "tcgen05.mma (UTC*MMA)": r"\bUTC\w*MMA\b",
    "tcgen05.ld/st (LDTM/STTM)": r"\b(LDTM|STTM)\b",
    "TMA bulk copy (UBLKCP/UTMALDG/UTMASTG)": r"\b(UBLKCP|UTMALDG|UTMASTG)\b",
    "legacy mma.sync (HMMA)": r"\bHMMA\b",
    "cp.async (LDGSTS)": r"\bLDGSTS\b",
    "DSMEM / cluster (ST.ASYNC... / UCGABAR / MAPA)": r"\b(STAS|UCGABAR_ARV|UCGABAR_WAIT|MAPA)\b",
    "mbarrier (SYNCS)": r"\bSYNCS\b",
    "packed fp32x2 (FADD2/FMUL2/FFMA2)": r"\b(FADD2|FMUL2|FFMA2)\b",
    "warp reduce (REDUX)": r"\bREDUX\b",
    "global reduction (RED / REDG)": r"\b(RED|REDG)\b",
}


def main():
    txt = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    out, name = {}, None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(anonymous namespace\)::", "", name).split("(")[0].replace("void ", "").replace("eda::", "")
            out[name] = {k: 0 for k in PAT}
            out[name]["instructions"] = 0
            continue
        if name and re.search(r"/\*[0-9a-f]{4,}\*/", line):
            out[name]["instructions"] += 1
            for k, p in PAT.items():
                if re.search(p, line):
                    out[name][k] += 1
    res = {k: {a: b for a, b in v.items() if b} for k, v in sorted(out.items())}
    js = json.dumps({"source": "cuobjdump -sass eda_b200/lib/libeda_b200.so (sm_100a)", "kernels": res}, indent=1)
    if len(sys.argv) > 1:
        open(sys.argv[1], "w").write(js)
    for k, v in res.items():
        tags = ", ".join(f"{a.split(' (')[0]}={b}" for a, b in v.items() if a != "instructions")
        print(f"{k[:70]:70s} {v.get('instructions', 0):6d}  {tags}")


if __name__ == "__main__":
    main()
