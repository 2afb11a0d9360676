#!/bin/bash
# A/B of prebuilt library variants on one GPU box, in one gpurun call (a call costs ~25 s of box time per variant).
#   make -C hana-softwarerenderer_b200/csrc OUT=$PWD/variants/<name>.so EXTRA=-D<SWITCH>      (one per variant, built on the CPU box)
#   gpurun --timeout 300 -- tools/ab_variants.sh base <name> ...
# Per variant: the bench's device-resident leg (10 steps, no CPU leg) and the GPU parity tests; the library that was in
# place is restored afterwards. Results: gpurun_out/ab_<name>.json, gpurun_out/ab_<name>_pytest.log.
mkdir -p gpurun_out
LIB=hana-softwarerenderer_b200/libhana_b200.so
cp $LIB /tmp/keep.so
for v in "$@"; do
  cp variants/$v.so $LIB || continue
  python bench.py --steps 10 --no-cpu-baseline > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/ab_$v.json").read().strip().splitlines()[-1])
    print("$v", round(d["value"]), "frames/s, ms/step %.3f" % d["ms_per_step"],
          {k: round(x, 3) for k, x in d["kernel_ms_per_step"].items()}, d["clocks"]["sm_mhz"], "MHz")
except Exception as e:
    print("$v FAILED", e)
PY
  timeout 200 python -m pytest tests -m gpu -x -q > gpurun_out/ab_${v}_pytest.log 2>&1
  echo "$v pytest: $(tail -1 gpurun_out/ab_${v}_pytest.log)"
done
cp /tmp/keep.so $LIB
