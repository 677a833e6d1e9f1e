# block width of the queued fan launch: best-of-6 time of one C2 subset
cd $GRAFT_REPO_ROOT
for bw in 4 8 16 32 64; do
  echo "== PAX_FAN_BW=$bw"
  PAX_FAN_BW=$bw timeout 300 python tools/profile_step.py --subsets 1 --reps 6 2>&1 | grep "rep [0-9]" | awk '{print $6}' | sort -n | head -3 | tr '\n' ' '; echo
done
PAX_FAN_BW=32 timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "fan" 2>&1 | tail -1
