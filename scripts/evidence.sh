# Developer helper: the per-round evidence pass on one B200 (bench lines of every workload, ncu launch list,
# ncu --set full captures of the dominant kernels).  usage: bash scripts/evidence.sh <tag>
# gpurun brings back at most 64 MiB of gpurun_out/: the captures are summarised ON THE BOX (ncu is there) and only
# the headline's report travels.
set -x
T=$1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_${T}.log 2>&1; tail -4 gpurun_out/smoke_${T}.log
python bench.py > gpurun_out/bench_${T}.json 2> gpurun_out/bench_${T}.err; tail -c 300 gpurun_out/bench_${T}.err
for w in cfg2_24mp_rgb8_linear cfg3_8k_rgba16_cubic cfg4_50mp_rgbf32_cubic cfg5_4k_rgb8_cubic cfg5_batch_4k_rgb8_cubic; do python bench.py --workload $w --no-cpu --steps 100 > gpurun_out/bench_${T}_$w.json 2> gpurun_out/bench_${T}_$w.err; done
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${T}_reference.json 2> gpurun_out/bench_${T}_reference.err
python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/plain_${T}.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${T}.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/nculist_${T}.log 2>&1
for w in fast rgba16 rgb8lin rgb8 none; do
	python scripts/profile_one.py $w 5 > gpurun_out/plain_${T}_$w.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:stream_kernel -s 2 -c 1 -f -o gpurun_out/prof_stream_${w}_${T} python scripts/profile_one.py $w 5 > gpurun_out/ncu_${T}_$w.log 2>&1
	python scripts/make_profile_md.py ncu "r01 ($T)" "ncu --set full of stream_kernel, profile_one.py $w" gpurun_out/prof_stream_${w}_${T}.ncu-rep gpurun_out/ncu_stream_${w}_${T}.md
	[ $w = fast ] || rm -f gpurun_out/prof_stream_${w}_${T}.ncu-rep
done
python scripts/profile_batch.py 32 > gpurun_out/plain_${T}_batch.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:stream_kernel -s 2 -c 1 -f -o gpurun_out/prof_stream_batch_${T} python scripts/profile_batch.py 32 > gpurun_out/ncu_${T}_batch.log 2>&1
python scripts/make_profile_md.py ncu "r01 ($T)" "ncu --set full of stream_kernel, ONE launch over 32 4K RGB8 frames, Cubic" gpurun_out/prof_stream_batch_${T}.ncu-rep gpurun_out/ncu_stream_batch_${T}.md
ncu -i gpurun_out/prof_stream_batch_${T}.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/source_stream_batch_${T}.csv.gz
rm -f gpurun_out/prof_stream_batch_${T}.ncu-rep
python scripts/quick_bench.py mix > gpurun_out/mix_${T}.log 2>&1
python scripts/quick_bench.py none >> gpurun_out/mix_${T}.log 2>&1
python scripts/quick_bench.py batch >> gpurun_out/mix_${T}.log 2>&1
python scripts/quick_bench.py target 2>&1 | grep -i exact >> gpurun_out/mix_${T}.log
python scripts/e2e_pageable.py >> gpurun_out/mix_${T}.log 2>&1
du -sh gpurun_out; ls gpurun_out | grep ${T} | wc -l
