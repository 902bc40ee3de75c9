set -x
python -m pytest tests/test_bank_gpu.py tests/test_channelizer_gpu.py -x -q -m gpu -k "fm or nbfm or airspy or golden" > gpurun_out/s4c_pytest.log 2>&1; tail -n 15 gpurun_out/s4c_pytest.log
CMD2="python bench.py --workload nbfm_4096 --steps 5 --warmup 3 --no-cpu-baseline --device-only"
$CMD2 > gpurun_out/s4c_nbfm.json 2> gpurun_out/s4c_nbfm.err; tail -c 1500 gpurun_out/s4c_nbfm.json
for t in 192 320 384; do SDRGPU_NBFM_TILE=$t $CMD2 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('tile $t', d['ms_per_step'])"; done
CMD3="python bench.py --workload nbfm_4096 --steps 1 --warmup 3 --no-cpu-baseline --device-only"
ncu --set full --clock-control none --import-source on -k regex:'nbfm' -s 3 -c 1 -o gpurun_out/s4c_prof_nbfm $CMD3 > gpurun_out/s4c_ncu.log 2>&1
tail -n 3 gpurun_out/s4c_ncu.log
