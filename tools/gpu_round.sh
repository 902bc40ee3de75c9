#!/bin/bash
# One GPU-box round: parity tests, smoke, bench (both arms), then the ncu launch list and the full capture of the
# dominant kernels.  Run as:  gpurun --timeout 1500 -- 'bash tools/gpu_round.sh r01'
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
cat $out/${tag}_bench.json
python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err
cat $out/${tag}_bench_reference.json
if [ "$2" != "noncu" ]; then
CMD="python bench.py --workload c4fm --steps 1 --warmup 3 --no-cpu-baseline --device-only"
$CMD > $out/${tag}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv $CMD > $out/${tag}_ncu1.log 2>&1
$CMD > $out/${tag}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'pfb|fir|psk' -s 9 -c 3 -o $out/${tag}_prof $CMD > $out/${tag}_ncu2.log 2>&1
tail -5 $out/${tag}_ncu2.log
fi
