#!/bin/bash
# first-contact GPU run: parity tests, then whatever the caller appends
set -o pipefail
python -m pytest tests -m gpu -x -q 2>&1 | tail -40
