#!/bin/bash
# A/B of prebuilt library variants on one GPU box, in one gpurun call (~20 s of box time per variant).
#   make -C hana-softwarerenderer_b200/csrc OUT=$PWD/variants/<name>.so EXTRA=-D<SWITCH>      (one per variant, built on the CPU box)
#   gpurun --timeout 600 -- tools/ab_variants.sh [--test] base <name> ...
# Per variant: the bench's device-resident leg (10 steps, parity check of every 64th frame, no CPU leg, no extras) and,
# with --test, the GPU parity tests; the library that was in place is restored afterwards.
# Results: gpurun_out/ab_<name>.json (+ gpurun_out/ab_<name>_pytest.log).
mkdir -p gpurun_out
LIB=hana-softwarerenderer_b200/libhana_b200.so
cp $LIB /tmp/keep.so
TEST=0
if [ "$1" = "--test" ]; then TEST=1; shift; fi
for v in "$@"; do
  cp variants/$v.so $LIB || continue
  python bench.py --steps 10 --no-cpu-baseline --no-extras > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/ab_$v.json").read().strip().splitlines()[-1])
    p = d["parity"]
    print("$v", round(d["value"]), "frames/s, ms/step %.3f" % d["ms_per_step"], "frac %.4f" % d["frame_roofline"]["frac"],
          {k: round(x, 3) for k, x in d["kernel_ms_per_step"].items()}, d["clocks"]["sm_mhz"], "MHz",
          "parity fail=%d cov=%d depth=%d cmax=%d cpx=%d" % (p["mismatches"], p["coverage_mismatch_px"], p["depth_bits_mismatch_px"], p["colour_maxdiff"], p["colour_mismatch_px"]),
          "e2e", round(d["e2e"]["value"]))
except Exception as e:
    print("$v FAILED", e, open("gpurun_out/ab_$v.err").read()[-400:])
PY
  if [ $TEST = 1 ]; then
    timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/ab_${v}_pytest.log 2>&1
    echo "$v pytest: $(tail -1 gpurun_out/ab_${v}_pytest.log)"
  fi
done
cp /tmp/keep.so $LIB
