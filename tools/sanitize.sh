#!/bin/bash
# compute-sanitizer (memcheck, racecheck, synccheck) over the frame parity tests of the CUDA path, on the GPU box:
#   gpurun --timeout 1800 -- tools/sanitize.sh r02
# Logs: gpurun_out/<tag>_sanitizer_{memcheck,racecheck,synccheck}.txt (copy to profiles/).
# Test selection: every kernel variant is exercised (all shaders, TMA and plain stores, RenderBuffer read-modify-write,
# ties, slivers, clip soup, band renders, async sweeps, the wide-slot shadow rasteriser, the micro-triangle
# visibility-buffer path on a scaled configs[3], the device TGA encoder); the full-size configs and the 1080p orbit are
# left out (racecheck runs 20-50x slower).
TAG=${1:-r02}
mkdir -p gpurun_out
SEL="not full_size and not every_64th and not dropin and not two_gpus and not split_frame and not african_head and not diablo and not render_tool and not c1_ and not orbit_files"
MEM="tests/test_gpu_parity.py tests/test_async_sweep.py tests/test_wide_slots.py tests/test_tga_rle.py tests/test_configs.py tests/test_sharding.py"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest $MEM -m gpu -x -q -k "$SEL" > gpurun_out/${TAG}_sanitizer_memcheck.txt 2>&1
echo "memcheck rc=$? $(tail -2 gpurun_out/${TAG}_sanitizer_memcheck.txt | tr '\n' ' ')"
RACE="tests/test_gpu_parity.py tests/test_wide_slots.py tests/test_tga_rle.py tests/test_configs.py"
SELR="$SEL and not c5_scaled and not soup and not copies"
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest $RACE -m gpu -x -q -k "$SELR" > gpurun_out/${TAG}_sanitizer_racecheck.txt 2>&1
echo "racecheck rc=$? $(tail -2 gpurun_out/${TAG}_sanitizer_racecheck.txt | tr '\n' ' ')"
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 1 python -m pytest $RACE -m gpu -x -q -k "$SELR" > gpurun_out/${TAG}_sanitizer_synccheck.txt 2>&1
echo "synccheck rc=$? $(tail -2 gpurun_out/${TAG}_sanitizer_synccheck.txt | tr '\n' ' ')"
