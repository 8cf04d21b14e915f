# test_gpu_wide.py under the cluster-kernel switches: 2 = tcgen05 forward only, 3 = tcgen05 backward only, 1 = both (default)
for m in ${MODES:-2 3 1}; do
  echo "=== IB200_CLUSTER_TC=$m"; IB200_CLUSTER_TC=$m timeout 600 python -m pytest tests/test_gpu_wide.py -m gpu -x -q -k "wide_encoder or full_length or wide_training" 2>&1 | tail -15
done
