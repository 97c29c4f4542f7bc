#!/bin/bash
# Regenerates tests/golden/matcher_reference_results.txt: what the reference's OWN src/matcher.cpp (compiled in place,
# oracle/_ref/libmatcherref.so) returns and writes (count + fingerprint of the pointers it stored) for every Matcher entry point on the fixed scenes of tests/tools/matcher_adapter_check.cpp
# (seed families 77 and 1000, 4 rounds each).  Needs /root/reference (or the prebuilt oracle/_ref) -- run from the repo root.
set -e
make -C oracle > /dev/null
g++ -std=c++11 -O1 -DREF_MATCHER -Ioracle/compat_myslam -Ioracle/compat -Iinclude -Itests/tools tests/tools/matcher_adapter_check.cpp \
    tests/tools/cabi_on_port.cpp -Loracle -lorbport -Loracle/_ref -lmatcherref -Wl,-rpath,$PWD/oracle -Wl,-rpath,$PWD/oracle/_ref -o /tmp/mcheck_real
{ for seed in 77 1000; do echo "# seed $seed"; /tmp/mcheck_real $seed | grep -E "^libm signature|^round|: [0-9]+ (matches|fused)"; done; } > tests/golden/matcher_reference_results.txt
wc -l tests/golden/matcher_reference_results.txt
