# Round-end evidence run (one B200): GPU tests, the bench line, the ncu launch list of the bench command, and
# ncu --set full captures of the dominant kernels (exported to text / CSV on the box: gpurun_out/ may carry <= 64 MiB back).
# usage (from the repo root): bash tools/run_round_profiles.sh <tag>
set -x
tag=${1:-r2}
o=gpurun_out
mkdir -p $o
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 > $o/${tag}_gputests.log
timeout 600 python __graft_entry__.py smoke 2>&1 | tail -3 > $o/${tag}_smoke.log; cat $o/${tag}_smoke.log
tail -3 $o/${tag}_gputests.log
timeout 900 python bench.py --steps 5 --warmup 3 > $o/${tag}_bench.json 2> $o/${tag}_bench.err
tail -c 400 $o/${tag}_bench.json
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $o/${tag}_bench_reference.json 2>> $o/${tag}_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $o/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train > $o/${tag}_ncu_bench.log 2>&1
export_rep() {  # <rep without extension>: details text + raw CSV, then drop the (large) report unless told to keep it
  ncu -i $1.ncu-rep --page details > $1_details.txt 2>/dev/null
  ncu -i $1.ncu-rep --page raw --csv > $1_raw.csv 2>/dev/null
  if [ "$2" != keep ]; then rm -f $1.ncu-rep; fi
}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:onet_decode -s 1 -c 1 -o $o/${tag}_onet_decode python tools/prof_decoder.py 1024 1 > $o/${tag}_ncu_dec.log 2>&1
export_rep $o/${tag}_onet_decode keep
timeout 900 ncu --set full --clock-control none -k regex:"mlp_chain_tc|fps_kernel|three_nn|onet_cbn|extract_mesh" -c 24 -o $o/${tag}_detection python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-train > $o/${tag}_ncu_det.log 2>&1
export_rep $o/${tag}_detection
timeout 600 ncu --set full --clock-control none -k regex:"query_and_group|grid_build|transpose_features" -c 12 -o $o/${tag}_qg python tools/prof_qg.py 4 1 > $o/${tag}_ncu_qg.log 2>&1
export_rep $o/${tag}_qg
timeout 300 python tools/prof_qg.py 4 30 > $o/${tag}_qg.log 2>&1
du -sh $o; ls -la $o | tail -20
