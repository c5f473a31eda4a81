# Developer helper: the per-round evidence pass on one B200 (bench lines of every workload, ncu launch list,
# ncu --set full captures of the dominant kernels).  usage: bash scripts/evidence.sh <tag>
set -x
T=$1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_${T}.log 2>&1; tail -4 gpurun_out/smoke_${T}.log
python bench.py > gpurun_out/bench_${T}.json 2> gpurun_out/bench_${T}.err; tail -c 300 gpurun_out/bench_${T}.err
for w in cfg2_24mp_rgb8_linear cfg3_8k_rgba16_cubic cfg4_50mp_rgbf32_cubic cfg5_4k_rgb8_cubic cfg5_batch_4k_rgb8_cubic; do python bench.py --workload $w --no-cpu --steps 100 > gpurun_out/bench_${T}_$w.json 2> gpurun_out/bench_${T}_$w.err; done
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${T}_reference.json 2> gpurun_out/bench_${T}_reference.err
python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/plain_${T}.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${T}.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/nculist_${T}.log 2>&1
for w in fast rgba16 rgb8lin rgb8 none; do python scripts/profile_one.py $w 5 > gpurun_out/plain_${T}_$w.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:stream_kernel -s 2 -c 1 -f -o gpurun_out/prof_stream_${w}_${T} python scripts/profile_one.py $w 5 > gpurun_out/ncu_${T}_$w.log 2>&1; done
python scripts/quick_bench.py mix > gpurun_out/mix_${T}.log 2>&1
python scripts/quick_bench.py none >> gpurun_out/mix_${T}.log 2>&1
python scripts/quick_bench.py batch >> gpurun_out/mix_${T}.log 2>&1
python scripts/quick_bench.py target 2>&1 | grep -i exact >> gpurun_out/mix_${T}.log
python scripts/e2e_pageable.py >> gpurun_out/mix_${T}.log 2>&1
ls gpurun_out | grep ${T} | wc -l
