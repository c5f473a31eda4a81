# Developer helper: the round-2 evidence pass on one B200 (bench lines, ncu launch list, ncu --set full captures of the
# dominant kernel of every bench workload -> profiles/r02_ncu_*.md and profiles/traffic.json).
#   usage (under gpurun): bash scripts/evidence_r02.sh <tag>
# gpurun brings back at most 64 MiB of gpurun_out/: the captures are summarised ON THE BOX (ncu is there) and only
# the headline's report travels.
set -x
T=$1
python bench.py > gpurun_out/bench_${T}.json 2> gpurun_out/bench_${T}.err; tail -c 300 gpurun_out/bench_${T}.err
python bench.py --workload cfg5_batch_4k_rgb8_cubic --no-cpu --steps 50 > gpurun_out/bench_${T}_cfg5_batch.json 2> gpurun_out/bench_${T}_cfg5_batch.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${T}_reference.json 2> gpurun_out/bench_${T}_reference.err
python bench.py --steps 2 --warmup 1 --no-cpu --no-workloads > gpurun_out/plain_${T}.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${T}.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-workloads > gpurun_out/nculist_${T}.log 2>&1
python scripts/make_profile_md.py list "r02 ($T)" "python bench.py --steps 2 --warmup 1 --no-cpu --no-workloads" gpurun_out/launches_${T}.csv gpurun_out/r02_launches_bench_${T}.md
for w in target_100mp_rgb16_cubic cfg2_24mp_rgb8_linear cfg3_8k_rgba16_cubic cfg4_50mp_rgbf32_cubic cfg5_4k_rgb8_cubic cfg5_batch_4k_rgb8_cubic target_100mp_rgb16_cubic:exact cfg2_24mp_rgb8_linear:exact cfg5_4k_rgb8_cubic:exact; do
	n=$(echo $w | tr ':' '_')
	python scripts/profile_one.py wl:$w 4 > gpurun_out/plain_${T}_$n.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'stream_kernel|tiled_kernel' -s 2 -c 1 -f -o gpurun_out/prof_${n}_${T} python scripts/profile_one.py wl:$w 4 > gpurun_out/ncu_${T}_$n.log 2>&1
	python scripts/make_profile_md.py ncu "r02 ($T)" "ncu --set full of the kernel bench.py workload $w launches (scripts/profile_one.py wl:$w; L2 flushed between launches)" gpurun_out/prof_${n}_${T}.ncu-rep gpurun_out/r02_ncu_${n}_${T}.md
	[ $w = target_100mp_rgb16_cubic ] || [ $w = cfg5_batch_4k_rgb8_cubic ] || rm -f gpurun_out/prof_${n}_${T}.ncu-rep
done
ncu -i gpurun_out/prof_cfg5_batch_4k_rgb8_cubic_${T}.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/source_cfg5_batch_${T}.csv.gz
rm -f gpurun_out/prof_cfg5_batch_4k_rgb8_cubic_${T}.ncu-rep
python scripts/quick_bench.py ab > gpurun_out/ab_${T}.log 2>&1
python scripts/quick_bench.py exact >> gpurun_out/ab_${T}.log 2>&1
python scripts/quick_bench.py exact8 >> gpurun_out/ab_${T}.log 2>&1
du -sh gpurun_out; ls gpurun_out | grep ${T} | wc -l
