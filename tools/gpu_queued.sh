# block width of the queued fan launch: best-of-6 time of one C2 subset; parity incl. the chord-split launches
cd $GRAFT_REPO_ROOT
for bw in 4 8 16 32; do
  echo "== PAX_FAN_BW=$bw"
  PAX_FAN_BW=$bw timeout 300 python tools/profile_step.py --subsets 1 --reps 6 2>&1 | grep "rep [0-9]" | awk '{print $6}' | sort -n | head -3 | tr '\n' ' '; echo
done
echo "== PAX_FAN_QUEUED=0"
PAX_FAN_QUEUED=0 timeout 300 python tools/profile_step.py --subsets 1 --reps 6 2>&1 | grep "rep [0-9]" | awk '{print $6}' | sort -n | head -3 | tr '\n' ' '; echo
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_gpu_fullsize_oracle.py -q -m "gpu and not slow" 2>&1 | tail -2
PAX_FAN_BW=5 timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "fan" 2>&1 | tail -1
