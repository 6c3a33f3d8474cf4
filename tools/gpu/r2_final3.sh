#!/bin/bash
# Last check at the final head (host-side changes after r2_final2): full GPU suite + smoke
O=gpurun_out/r2_final3; mkdir -p $O
s=$(date +%s); timeout 1200 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$? wall=$(( $(date +%s)-s ))s"; tail -3 $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"
