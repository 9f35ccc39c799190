timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
bash scripts/gpu_variants.sh
bash scripts/gpu_profile.sh self-supervised-normal-depth-estimation_b200/variants/libsnde_p2c3.so
