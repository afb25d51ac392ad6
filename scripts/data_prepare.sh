#!/bin/bash
# Parity: reference src/data_prepare.sh
python -m atomo_b200.data.data_prepare "$@"
