#!/usr/bin/env python
"""Turns the MARGIN lines of the full-size parity tests into profiles/rN_parity_margins.md.

  A2C_TEST_REPORT=1 python -m pytest tests/test_gpu_parity.py -q -s -k full_size | grep MARGIN > margins.txt
  python profiles/margins.py margins.txt profiles/r2_parity_margins.md
"""
import re
import sys


def main(src, out):
    pat = re.compile(r"MARGIN (\S+) prec=(\d) (\S+)\s+numel=(\d+)\s+e_full=(\S+) e_sampled=(\S+) "
                     r"ref_own=(\S+) bar=(\S+) used=(\S+) (\S+)")
    groups = {}
    for line in open(src):
        m = pat.search(line)
        if m:
            f, prec, name, numel, ef, es, ro, bar, used, clause = m.groups()
            groups.setdefault((f, int(prec)), []).append(
                (name, int(numel), float(ef), float(es), float(ro), float(bar), float(used), clause))
    L = ["# Gradient parity margins at full size (C2, C3, C4), measured on B200", "",
         "One row per parameter tensor of `tests/test_gpu_parity.py::test_full_size_update_against_reference_golden`.",
         "`e_full` = max over EVERY element of |cuda - float64 oracle| / max(max|float64|, 1e-3 x the model's largest",
         "gradient entry); the float64 oracle is evaluated on the GPU's fp64 units and first pinned (2e-7) to the",
         "committed CPU-generated float64 fixture.  `e_sampled` = the same against the fixture's norm + 4,096 strided",
         "samples.  `ref_own` = the unmodified fp32 reference's own distance from the float64 values (same metric,",
         "from the fixtures).  Bar: `max(1e-5, ref_own) + k * ref_own` with k = 2 (precision 0, fp32 FFMA) and k = 8",
         "(precision 1, 3xTF32 on tcgen05); `used` = e_full / bar as printed by the run that produced this file",
         "(its bar may have used an earlier k; the ratios in the section titles are what k was chosen from);",
         "`clause` = which term of the bar decides.", ""]
    for (f, prec), rows in sorted(groups.items()):
        ratios = [r[2] / r[4] for r in rows if r[4] > 0 and r[2] > 1e-5]
        L.append("## %s, precision %d  (max e_full / ref_own among tensors above 1e-5: %.2f)"
                 % (f, prec, max(ratios) if ratios else 0.0))
        L.append("")
        L.append("| tensor | numel | e_full | e_sampled | ref_own | bar | used | clause |")
        L.append("|---|---|---|---|---|---|---|---|")
        for name, numel, ef, es, ro, bar, used, clause in rows:
            L.append("| %s | %d | %.3g | %.3g | %.3g | %.3g | %.2f | %s |" % (name, numel, ef, es, ro, bar, used, clause))
        L.append("")
    open(out, "w").write("\n".join(L) + "\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
