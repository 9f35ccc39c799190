timeout 600 python -m pytest tests -m gpu -q -k "identity_pose" 2>&1 | grep -E "^E|passed|failed" | head -20
