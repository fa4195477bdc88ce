set -x
ncu --set full --clock-control none --import-source on -k regex:"pyr_fast|octree|describe|hamming_stream" -s 33 -c 11 -f -o gpurun_out/r1_final3 python bench.py --steps 1 --warmup 3 --frames 32 --no-cpu-baseline > gpurun_out/ncu_final.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"bow_descend|group_distance" -s 4 -c 3 -f -o gpurun_out/r1_bow python tools/bow_profile.py > gpurun_out/ncu_bow.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_final3.csv python bench.py --steps 2 --warmup 3 --frames 128 --no-cpu-baseline > gpurun_out/b.log 2>&1
python tools/bow_profile.py > gpurun_out/bow_timing.log 2>&1; tail -1 gpurun_out/bow_timing.log
python bench.py > gpurun_out/bench_r1_final.json 2> gpurun_out/bench_r1_final.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r1_reference.json 2>> gpurun_out/bench_r1_final.err
tail -c 1500 gpurun_out/bench_r1_final.json; tail -c 600 gpurun_out/bench_r1_reference.json
