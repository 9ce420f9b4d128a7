#!/bin/bash
# first GPU contact: tests + smoke
set -x
nvidia-smi --query-gpu=name,memory.total --format=csv
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -20
python -m pytest tests -m gpu -q 2>&1 | tail -40
