#!/bin/bash
# One GPU-box round: parity tests, smoke, bench (both arms), then the ncu launch
# list and the full capture of the dominant kernels.  Run as:  gpurun --timeout 1800 -- 'bash tools/gpu_round.sh r02'
# ncu only runs after the identical plain command exited 0 (B200_PROFILING.md).
tag=${1:-rXX}
out=gpurun_out
mkdir -p $out
set -x
python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $out/${tag}_pytest.log
tail -3 $out/${tag}_pytest.log
python __graft_entry__.py smoke > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $out/${tag}_smoke.log
tail -2 $out/${tag}_smoke.log
python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"
head -c 1500 $out/${tag}_bench.json; echo
python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err
head -c 600 $out/${tag}_bench_reference.json; echo
# (compute-sanitizer memcheck / racecheck of the smoke path were part of this round until the pool closed the tool:
#  "compute-sanitizer is closed on this pool and stays closed: runs under it have left GPUs needing a reset" -- r2j_memcheck.log.
#  Bounds are covered by the parity tests instead: ragged / oversized / empty calls, odd channel counts, partial tiles.)
# side measurements behind DESIGN.md section 4 / 5 (seconds each): FIR stage alone, blocking against streamed e2e step
python tools/fir_time.py 6400 > $out/${tag}_fir_time.txt 2>&1
python tools/e2e_probe.py 8 8 > $out/${tag}_e2e_probe.txt 2>&1
if [ "$2" != "noncu" ]; then
CMD="python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline --device-only"
$CMD > $out/${tag}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'pfb|fir_agc|psk|carry|append|convert|osc|halfband|fm_|nbfm|save_state' -c 1200 --csv --log-file $out/${tag}_launches.csv $CMD > $out/${tag}_ncu1.log 2>&1
$CMD > $out/${tag}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'pfb2|fir_agc|psk' -s 30 -c 12 -o $out/${tag}_prof $CMD > $out/${tag}_ncu2.log 2>&1
tail -5 $out/${tag}_ncu2.log
CMD2="python bench.py --workload nbfm_4096 --steps 1 --warmup 3 --no-cpu-baseline --device-only"
$CMD2 > $out/${tag}_plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'nbfm' -s 3 -c 1 -o $out/${tag}_prof_nbfm $CMD2 > $out/${tag}_ncu3.log 2>&1
fi
