#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
# x0 in {-4,-2,3,2,1276}, y0 = 4 (flag 2) or -2 (flag 0)
for x in -4 -2 3 2 1276; do for f in 2 0; do echo -n "x0=$x yflag=$f  "; scripts/ubench/tma_test 1 $((x*16+f)); done; done
