#!/bin/bash
# Parity: reference tools/install.sh (pdsh + key distribution).  Pushes the repo to every node.
set -e
HOSTS=${1:-tools/hosts_address}
for h in $(cat "$HOSTS"); do
  rsync -az --exclude .git --exclude gpurun_out ./ "$h":~/atomo_b200/ &
done
wait
python tools/cluster.py run_command "cd ~/atomo_b200 && bash tools/pre_run.sh" --hostfile "$HOSTS"
