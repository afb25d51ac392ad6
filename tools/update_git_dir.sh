#!/bin/bash
# Parity: reference tools/update_git_dir.sh — update the checkout on every node.
python tools/cluster.py run_command "cd ~/atomo_b200 && git pull --ff-only && python setup.py build_ext --inplace"
