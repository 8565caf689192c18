#!/bin/bash
# The command sequences of the `gpurun` calls whose outputs are summarised under profiles/.
#   gpurun --timeout 1500 -- 'bash tools/gpu.sh TAG step [step ...]'
# steps (each writes gpurun_out/TAG_*):
#   test            pytest -m gpu (whole suite)
#   bench           python bench.py (default flags: the driver's command)
#   benchlong       bench.py --steps 300 --warmup 20 --no-cpu-baseline
#   ref             bench.py --impl reference (default flags)
#   launches        ncu launch list of bench.py --steps 3 --warmup 3
#   full            ncu --set full of one whole sweep (NCU_SKIP / NCU_COUNT select it: 8 kernels -> NCU_SKIP=40 NCU_COUNT=8) + raw csv
#   traffic         regenerates profiles/ncu_traffic.json from the full capture of this run (stamped with the tree's hash)
#   sanitize        compute-sanitizer memcheck + racecheck over the reduced selection (tools/sanitize_cases.py)
#   sanitize_mgpu   memcheck over the reduced 2-GPU segment-split worker, both transports (needs --gpus 2)
#   mgpu:N          multi-GPU parity workers (tests/test_multi_gpu.py) + bench --gpus N on the N GPUs of this box
#   latency         tools/scan_latency.py (C1, C2, C4/8 stage tables)
TAG=$1; shift
mkdir -p gpurun_out
OUT=gpurun_out/$TAG
nvidia-smi --query-gpu=index,name,clocks.max.sm --format=csv > ${OUT}_gpu.txt
for step in "$@"; do
  case $step in
    test)
      timeout 1500 python -m pytest tests -m gpu -x -q > ${OUT}_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -6 ${OUT}_pytest_gpu.log ;;
    bench)
      timeout 900 python bench.py > ${OUT}_bench_line.json 2> ${OUT}_bench.err; echo "bench exit $?"; tail -2 ${OUT}_bench.err
      python tools/show_line.py ${OUT}_bench_line.json ;;
    benchlong)
      timeout 600 python bench.py --steps 300 --warmup 20 --no-cpu-baseline > ${OUT}_bench_long_line.json 2> ${OUT}_bench_long.err; echo "bench exit $?"
      python tools/show_line.py ${OUT}_bench_long_line.json ;;
    ref)
      timeout 1200 python bench.py --impl reference > ${OUT}_bench_reference_line.json 2> ${OUT}_bench_ref.err; echo "ref exit $?"
      head -c 1500 ${OUT}_bench_reference_line.json; echo ;;
    launches)
      timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file ${OUT}_launches_bench_steps3.csv \
        python bench.py --steps 3 --warmup 3 --no-cpu-baseline > ${OUT}_launches.out 2>&1; echo "ncu list exit $?" ;;
    full)
      timeout 900 ncu --set full --clock-control none --import-source on \
        -k regex:"k_cand_|k_block_emit|k_fwd_|k_bwd_|k_reduce_" -s ${NCU_SKIP:-56} -c ${NCU_COUNT:-13} -f -o ${OUT}_full \
        python bench.py --steps 3 --warmup 3 --no-cpu-baseline > ${OUT}_full.out 2>&1; echo "ncu full exit $?"
      ncu -i ${OUT}_full.ncu-rep --page raw --csv > ${OUT}_sweep_kernels_ncu_full_raw.csv 2>/dev/null ;;
    traffic)
      python tools/ncu_traffic.py ${OUT}_sweep_kernels_ncu_full_raw.csv ${OUT}_ncu_traffic.json ;;
    sanitize)
      for tool in memcheck racecheck; do
        timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 python tools/sanitize_cases.py > ${OUT}_sanitizer_$tool.log 2>&1
        echo "$tool exit $?"; tail -4 ${OUT}_sanitizer_$tool.log
      done ;;
    sanitize_mgpu)
      # the reduced segment-split worker under memcheck, both transports (peer mailboxes over NVLink, NCCL all-gathers)
      for tr in peer nccl; do
        HML_MGPU_SMALL=1 HML_EXCHANGE=$tr HML_EXPECT_TRANSPORT=$tr timeout 900 compute-sanitizer --tool memcheck --target-processes all \
          --error-exitcode 9 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 \
          tests/mgpu_worker.py > ${OUT}_sanitizer_memcheck_2gpu_$tr.log 2>&1
        echo "memcheck 2gpu $tr exit $?"; grep -E "ERROR SUMMARY|MGPU WORKER OK" ${OUT}_sanitizer_memcheck_2gpu_$tr.log | tail -4
      done ;;
    mgpu:*)
      N=${step#mgpu:}
      HML_TEST_WORLDS=$N timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > ${OUT}_pytest_mgpu_n$N.log 2>&1; echo "mgpu pytest exit $?"; tail -6 ${OUT}_pytest_mgpu_n$N.log
      timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
        bench.py --gpus $N > ${OUT}_bench_${N}gpu_line.json 2> ${OUT}_bench_${N}gpu.err; echo "bench N=$N exit $?"
      python tools/show_line.py ${OUT}_bench_${N}gpu_line.json ;;
    latency)
      timeout 600 python tools/scan_latency.py > ${OUT}_scan_latency.json 2> ${OUT}_scan_latency.err; echo "latency exit $?"; cat ${OUT}_scan_latency.json ;;
    *) echo "unknown step $step" ;;
  esac
done
