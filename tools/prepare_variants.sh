#!/bin/bash
# Builds the prepared experiment variants of the library next to the default one (run HERE, before a gpurun call; ~5 min each):
#   regorder  traversal orders the surviving children in registers (no local-memory arrays)          -DZR_TRAVERSE_REGISTER_ORDER
#   skipzero  zero-direction emissive-hit queries stopped in FindClosestEmissive                       -DZR_SKIP_ZERO_WI_QUERIES
#   rgicut    rgi.cu built WITH the traversal-level degenerate-ray cut (the measured-slow k_rgi)       -DZR_RGI_WITH_DEGENERATE_CUT
# All three are host-verified for correctness (tests/test_bvh_random.py, tests/test_device_source_vs_oracle.py); none is timed yet.
set -e
cd "$(dirname "$0")/.."
vars=("$@"); [ ${#vars[@]} -eq 0 ] && vars=(regorder skipzero rgicut)
for v in "${vars[@]}"; do
  case $v in
    regorder) f=-DZR_TRAVERSE_REGISTER_ORDER ;;
    skipzero) f=-DZR_SKIP_ZERO_WI_QUERIES ;;
    rgicut)   f=-DZR_RGI_WITH_DEGENERATE_CUT ;;
    *) echo "unknown variant $v"; exit 1 ;;
  esac
  ZR_VARIANT=$v ZR_EXTRA_FLAGS=$f python -m zetaray_b200.build
  grep -q "zetaray_b200/build_$v/" .gpurunignore || echo "zetaray_b200/build_$v/" >> .gpurunignore
done
