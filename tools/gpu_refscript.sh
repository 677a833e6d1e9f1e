#!/bin/bash
# Run the test that executes the reference's OWN script (pCT_CBCT_Reconstruction.py, unmodified) on a GPU box.
# /root/reference does not travel with the repo snapshot and its sources are never copied into the repo, so the
# script is shipped inside the job's command line and unpacked to /tmp on the box:
#   gpurun -- bash -c "$(tools/gpu_refscript.sh --emit)"
REF=/root/reference/Physics-based-Augmentation/OSSART-Reconstruction/pCT_CBCT_Reconstruction.py
if [ "$1" = "--emit" ]; then
  echo "mkdir -p /tmp/refscript gpurun_out && echo $(base64 -w0 $REF) | base64 -d > /tmp/refscript/pCT_CBCT_Reconstruction.py && sha256sum /tmp/refscript/pCT_CBCT_Reconstruction.py && PAX_REFERENCE_SCRIPT=/tmp/refscript/pCT_CBCT_Reconstruction.py timeout 900 python -m pytest tests/test_io_and_cases.py -q -m gpu -k reference_script -s 2>&1 | tail -15 | tee gpurun_out/refscript_test.log"
fi
