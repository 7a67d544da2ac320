set -x
mkdir -p gpurun_out
for rep in 1 2 3; do
  for v in old new; do
    if [ $v = old ]; then d=_ab_old; else d=.; fi
    (cd $d && timeout 300 python tools/prof_decoder.py 256 5 fp16 2 2>&1 | grep decode | sed "s/^/$v /")
    (cd $d && timeout 300 python tools/prof_decoder.py 1024 3 fp16 2 2>&1 | grep decode | sed "s/^/$v /")
  done
done 2>&1 | grep -v "^+" | tee gpurun_out/ab_log.txt
