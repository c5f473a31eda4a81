set -x
T=$1
python bench.py > gpurun_out/bench_${T}.json 2> gpurun_out/bench_${T}.err; tail -c 600 gpurun_out/bench_${T}.err
for w in cfg2_24mp_rgb8_linear cfg3_8k_rgba16_cubic cfg4_50mp_rgbf32_cubic cfg5_4k_rgb8_cubic; do python bench.py --workload $w --no-cpu --steps 100 > gpurun_out/bench_${T}_$w.json 2> gpurun_out/bench_${T}_$w.err; done
python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/plain_${T}.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${T}.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/nculist_${T}.log 2>&1
for w in fast rgba16 rgb8lin rgb8; do python scripts/profile_one.py $w 5 > gpurun_out/plain_${T}_$w.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:stream_kernel -s 2 -c 1 -f -o gpurun_out/prof_stream_${w}_${T} python scripts/profile_one.py $w 5 > gpurun_out/ncu_${T}_$w.log 2>&1; done
ls -la gpurun_out | tail -20
