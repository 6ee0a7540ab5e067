mkdir -p gpurun_out
for wl in bench_qft bench_random_tn; do
  for mini in time flops; do
    echo "== $wl minimize=$mini"
    WL=$wl timeout 600 python scripts/searched_path_probe.py 256 $mini 2>&1 | cut -c1-420 | grep -v "^amplitude"
  done
done > gpurun_out/searched_other_configs.txt 2>&1
cat gpurun_out/searched_other_configs.txt
