python -m pytest tests/test_gpu_parity.py -q -m gpu -k "bulk_copy or backward_variants or fused_linearize_backward" 2>&1 | tail -15
for B in 131072 32768 4096; do for bulk in 0 1; do
echo "#### B=$B ILQR_FUSED=0 ILQR_BACKWARD_BULK=$bulk"
QG_ITERS=4 ILQR_FUSED=0 ILQR_BACKWARD_LANES=0 ILQR_BACKWARD_BULK=$bulk python scripts/quick_gpu.py $B 500 rk4 2>&1 | grep -i "per iteration\|solve"
done; done
