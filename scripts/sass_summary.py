"""Per-kernel SASS mnemonic counts of the in-tree libtvc.so -> profiles/sass_summary.md.

Evidence that kernel (a) is tcgen05 / TMEM / TMA code (UTCHMMA*, UTMALDG*, LDTM, UTCBAR), that kernel (b)'s
producers use bulk copies (UBLKCP), that kernels (b)/(c) move data with 128-bit accesses (LDG.E.128 /
STG.E.128 / LDS.128) and that kernel (c) issues reductions (RED / ATOMS) behind a warp match (MATCH.ANY).
Runs on the build container (cuobjdump only, no GPU).  usage: python scripts/sass_summary.py
"""
import collections
import hashlib
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
LIB = ROOT / "multimodal_detection_consistency_b200" / "libtvc.so"
OUT = ROOT / "profiles" / "sass_summary.md"

# column -> regex on the mnemonic (with modifiers) of one SASS instruction
COLUMNS = [
    ("UTCHMMA", r"^UTCHMMA(?!\.2CTA)"),
    ("UTCHMMA.2CTA", r"^UTCHMMA\.2CTA"),
    ("UTMALDG", r"^UTMALDG"),
    ("UTCBAR", r"^UTCBAR"),
    ("LDTM", r"^LDTM"),
    ("UBLKCP", r"^UBLKCP"),
    ("SYNCS", r"^SYNCS"),
    ("LDG.128", r"^LDG\..*128"),
    ("STG.128", r"^STG\..*128"),
    ("LDS.128", r"^LDS\..*128"),
    ("RED", r"^RED"),
    ("ATOMS", r"^ATOMS"),
    ("MATCH", r"^MATCH"),
    ("SHFL", r"^SHFL"),
    ("FMNMX3", r"^FMNMX3"),
    ("DFMA", r"^DFMA"),
    ("FFMA", r"^FFMA"),
    ("local LD/ST", r"^(LDL|STL)"),
]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    short = []
    for n in out:
        n = re.sub(r"\(anonymous namespace\)::", "", n)
        n = re.sub(r"^void ", "", n)
        n = re.sub(r"\(.*$", "", n)           # drop the parameter list
        short.append(n.replace("tvc::", ""))
    return short


def main():
    if not LIB.exists():
        sys.exit("build libtvc.so first (python -m multimodal_detection_consistency_b200.build)")
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    ins = re.compile(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)")
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = ins.match(line)
        if m and cur is not None:
            cur["_total"] += 1
            for col, rx in COLUMNS:
                if re.match(rx, m.group(1)):
                    cur[col] += 1
    names = demangle(list(kernels))
    used = [c for c, _ in COLUMNS if any(k[c] for k in kernels.values())]
    digest = hashlib.sha256(LIB.read_bytes()).hexdigest()[:16]
    lines = [
        "# SASS summary of `libtvc.so` (sm_100a)",
        "",
        f"`cuobjdump -sass multimodal_detection_consistency_b200/libtvc.so`, library sha256 `{digest}…`, "
        f"{len(kernels)} kernels; written by `scripts/sass_summary.py`.  Counts are static instructions per kernel.",
        "",
        "What to read off: `UTCHMMA(.2CTA)` = `tcgen05.mma` (cta_group::1 / ::2), `UTMALDG` = TMA tensor load "
        "(`cp.async.bulk.tensor`), `UTCBAR` = `tcgen05.commit` (mbarrier arrive, multicast in the pair kernel), "
        "`LDTM` = `tcgen05.ld` (TMEM -> registers), `UBLKCP` = `cp.async.bulk` (kernel (b) producers, row gathers "
        "from own / peer HBM), `SYNCS` = mbarrier operations, `MATCH` = `match.any` (warp aggregation before "
        "`RED` in kernel (c), duplicate detection in kernel (b)), `DFMA` = the fp64 statistics of kernel (b).",
        "",
        "| kernel | instr | " + " | ".join(used) + " |",
        "|---|---:|" + "---:|" * len(used),
    ]
    for (mangled, cnt), name in sorted(zip(kernels.items(), names), key=lambda t: t[1]):
        lines.append(f"| `{name}` | {cnt['_total']} | " + " | ".join(str(cnt[c]) if cnt[c] else "·" for c in used) + " |")
    OUT.write_text("\n".join(lines) + "\n")
    print(f"wrote {OUT} ({len(kernels)} kernels)")


if __name__ == "__main__":
    main()
