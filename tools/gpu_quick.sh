cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_parity.py -q -m gpu -k "fan" 2>&1 | tail -2
python tools/profile_step.py --subsets 1 --reps 6 2>&1 | grep "rep [0-9]" | awk '{print $6}' | sort -n | head -3 | tr '\n' ' '; echo
ncu --metrics l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:k_ax_fan -s 1 -c 1 python tools/profile_step.py --subsets 1 --reps 1 2>&1 | grep -E "l1tex__|lts__|gpu__time|smsp__inst"
