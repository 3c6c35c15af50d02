#!/bin/bash
# ncu evidence of a round (run on the GPU box through gpurun): the launch list of the bench command and a full
# capture of the kernels of the path, exported as csv next to it.  Usage: bash tools/profile_round.sh r02x
set -u
tag=${1:-r02}
out=gpurun_out/$tag
mkdir -p $out
BENCH="python bench.py --steps 2 --warmup 3 --chunks-per-step 2 --no-cpu-baseline --no-e2e --no-other-configs --file-pairs 0"
$BENCH > $out/bench_plain.json 2> $out/bench_plain.err || { echo "plain run failed"; tail -5 $out/bench_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 40 -c 42 --csv --log-file $out/launches.csv $BENCH > $out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_emit|k_match_plan|k_scan_newlines|k_gather_newlines' -s 16 -c 4 -f -o $out/prof_path $BENCH > $out/ncu_full.log 2>&1
for k in k_emit k_match_plan k_scan_newlines k_gather_newlines; do
  ncu -i $out/prof_path.ncu-rep --page raw --csv -k regex:$k > $out/raw_$k.csv 2>/dev/null
done
ncu -i $out/prof_path.ncu-rep --page source --print-source sass --csv -k regex:k_emit > $out/sass_k_emit.csv 2>/dev/null
# the gzip kernels: through the C ABI with MGK_OUT_GZIP (tools/gzip_timing.py)
python tools/gzip_timing.py > $out/gzip_plain.txt 2>&1 && \
ncu --set full --clock-control none -k regex:'k_gz_blocks' -s 2 -c 1 -f -o $out/prof_gz python tools/gzip_timing.py > $out/ncu_gz.log 2>&1
ncu -i $out/prof_gz.ncu-rep --page raw --csv > $out/raw_k_gz_blocks.csv 2>/dev/null
rm -f $out/prof_gz.ncu-rep
ls -la $out
