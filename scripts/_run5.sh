python scripts/host_costs.py 2>&1 | tail -6
python scripts/step_profile.py --steps 30 2>&1 | grep -v Warning | head -48 | cut -c1-200
python scripts/pyprofile_step.py voc_scribble_b1 300 2>&1 | head -45 | cut -c1-160
