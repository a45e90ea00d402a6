# Full single-GPU check of a frozen build (run under gpurun): new-module tests first, whole GPU suite, default bench, smoke.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_vae.py -m gpu -q -s -k "encode or gaussian" 2>&1 | grep -E "parity|passed|failed|Error|assert" | tail -40 > gpurun_out/final_encoder.txt
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/final_pytest.txt
timeout 600 python bench.py > gpurun_out/final_bench.txt 2> gpurun_out/final_bench.err
timeout 300 python __graft_entry__.py smoke > gpurun_out/final_smoke.txt 2>&1
cat gpurun_out/final_encoder.txt; tail -4 gpurun_out/final_pytest.txt; grep '^{' gpurun_out/final_bench.txt | cut -c1-200; tail -2 gpurun_out/final_smoke.txt
