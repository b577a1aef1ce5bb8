set -x
O=gpurun_out/final
mkdir -p $O
python bench.py > $O/n1.json 2> $O/n1.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/ref.json 2> $O/ref.err
for wl in mimc_2p14 aggregation_16; do python bench.py --workload $wl --steps 20 --warmup 5 --inflight 1 > $O/$wl.json 2>/dev/null; done
python bench.py --workload training_8192 --steps 10 --warmup 3 --inflight 8 --no-cpu-baseline > $O/training_8192_x8.json 2>/dev/null
python bench.py --workload training_8192 --steps 10 --warmup 3 --inflight 1 --no-cpu-baseline > $O/training_8192.json 2>/dev/null
for wl in mimc_2p20 training_2p20 mimc_2p22; do python bench.py --workload $wl --steps 3 --warmup 3 --inflight 1 --no-cpu-baseline > $O/$wl.json 2>/dev/null; done
python tools/prove_once.py training_2p16 2 > $O/prove_once.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python tools/prove_once.py training_2p16 2 > $O/ncu1.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'k_ntt_pass|k_transpose_cols' -c 60 --csv --log-file $O/k1k2_traffic.csv python tools/prove_once.py training_2p16 2 > $O/ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_ntt_pass -s 8 -c 2 -o $O/ntt_final python tools/prove_once.py training_2p16 2 > $O/ncu3.log 2>&1
ls -la $O
