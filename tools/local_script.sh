#!/bin/bash
# Parity: reference tools/local_script.sh — run from the launch node: refresh host files, push code.
python tools/cluster.py get_hosts "$@"
bash tools/install.sh tools/hosts_address
