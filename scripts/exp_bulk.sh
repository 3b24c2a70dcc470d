# bulk-copy ring vs per-thread cp.async ring of the thread-per-trajectory K2 (two-kernel path): parity tests, then timings
python -m pytest tests/test_gpu_parity.py tests/test_user_system.py -q -m gpu -k "bulk_copy or backward_variants" 2>&1 | tail -3
for B in 131072 32768 4096; do for bulk in 0 1; do
echo "#### B=$B ILQR_FUSED=0 ILQR_BACKWARD_BULK=$bulk"
QG_ITERS=4 ILQR_FUSED=0 ILQR_BACKWARD_LANES=0 ILQR_BACKWARD_BULK=$bulk python scripts/quick_gpu.py $B 500 rk4 2>&1 | grep -i "per iteration"
done; done
