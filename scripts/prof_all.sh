set -x
for w in persistent blocks batch_run batch_solve; do
  python scripts/prof_target.py $w --write-expected > gpurun_out/r02_prof_$w.log 2>&1 || { echo "FAILED plain $w"; tail -5 gpurun_out/r02_prof_$w.log; continue; }
  case $w in persistent) K=k_pdhg_persistent;; blocks) K=k_pdhg_blocks;; batch_run) K=k_batch_run;; batch_solve) K=k_batch_solve;; esac
  timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$K -c 1 -f -o gpurun_out/r02_$w python scripts/prof_target.py $w > gpurun_out/r02_ncu_$w.log 2>&1 || { echo "FAILED ncu $w"; tail -5 gpurun_out/r02_ncu_$w.log; }
done
ls -la gpurun_out/*.ncu-rep
