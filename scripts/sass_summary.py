"""profiles/sass_summary.txt: which Blackwell / NVSwitch instructions the built library actually contains.

    python scripts/sass_summary.py            (needs cuobjdump; runs on the build container, no GPU)

Counts SASS mnemonics per kernel family in prob_matrix_factorization_b200/libpmf_b200.so:
  UTCHMMA  tcgen05.mma            LDTM    tcgen05.ld (TMEM -> registers)     UTCBAR  tcgen05.commit -> mbarrier
  UBLKCP   cp.async.bulk (TMA)    SYNCS   mbarrier try_wait / arrive         FFMA2 / FADD2  packed FP32 pairs
  LDGMC    multimem.ld_reduce (in-switch reduction)     STG...STRONG.SYS next to it: multimem.st / release stores
  REDG / ATOMG  global reductions / atomics             MEMBAR  fences
"""
import collections
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(REPO, "prob_matrix_factorization_b200", "libpmf_b200.so")
MNEMONICS = ["UTCHMMA", "LDTM", "UTCBAR", "UBLKCP", "SYNCS", "FFMA2", "FADD2", "LDGMC", "STRONG.SYS", "REDG", "ATOMG", "MEMBAR",
             "DFMA", "MUFU"]
FAMILIES = [("gamma_pass_kernel", "gamma_pass"), ("gamma_multi_kernel", "gamma_multi"), ("gamma_combine_kernel", "gamma_combine"),
            ("topn_filter_kernel", "topn_filter (tcgen05)"), ("topn_", "topn other"), ("gauss_", "gauss"), ("lazy_step_kernel", "lazy_step"),
            ("hpf_map_", "hpf_map other"), ("lazy_", "lazy other"), ("adam_", "adam"), ("radix_", "radix sort"), ("elbo_", "elbo"),
            ("digamma_pass_kernel", "digamma_pass"), ("eval_stats_kernel", "eval_stats"), ("loop_decide_kernel", "loop_decide")]


def family(name):
    for key, fam in FAMILIES:
        if key in name:
            return fam
    return "other"


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts = collections.defaultdict(collections.Counter)
    kernels = collections.Counter()
    fam = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fam = family(m.group(1))
            kernels[fam] += 1
            continue
        if fam is None or "/*" not in line:
            continue
        for mn in MNEMONICS:
            if mn in line:
                counts[fam][mn] += 1
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
    out = [f"SASS summary of {os.path.relpath(LIB, REPO)}  (cuobjdump -sass; arch {', '.join(arch)})", __doc__.split("Counts SASS")[1].rstrip(), "",
           f"{'kernel family':28s} {'kernels':>7s} " + " ".join(f"{m:>10s}" for m in MNEMONICS)]
    for fam in sorted(kernels):
        out.append(f"{fam:28s} {kernels[fam]:7d} " + " ".join(f"{counts[fam][m]:10d}" for m in MNEMONICS))
    tot = collections.Counter()
    for c in counts.values():
        tot.update(c)
    out.append(f"{'TOTAL':28s} {sum(kernels.values()):7d} " + " ".join(f"{tot[m]:10d}" for m in MNEMONICS))
    text = "\n".join(out) + "\n"
    path = os.path.join(REPO, "profiles", "sass_summary.txt")
    with open(path, "w") as f:
        f.write(text)
    sys.stdout.write(text)


if __name__ == "__main__":
    main()
