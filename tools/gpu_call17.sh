mkdir -p gpurun_out
./tools/dmma_loop_bench > gpurun_out/dmma_loop_bench.txt 2>&1; echo rc=$?
cat gpurun_out/dmma_loop_bench.txt
