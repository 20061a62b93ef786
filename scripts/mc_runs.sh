for mc in 1 0; do echo "MLLP_ROWPART_MC=$mc"; MLLP_ROWPART_MC=$mc timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29537 scripts/rowpart_bench.py --quick $2 > gpurun_out/r02_mc_n$1_$mc.log 2>&1; echo EXIT $?; grep ROWPART_JSON gpurun_out/r02_mc_n$1_$mc.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l[len('ROWPART_JSON '):])
    for k,v in d.items(): print(k, 'mc' if v['multicast'] else 'p2p', round(v['us_per_iteration'],2), v['timeline_us'], v['parity_vs_oracle_K100'])
"; tail -3 gpurun_out/r02_mc_n$1_$mc.log | cut -c1-300; done
