#!/bin/bash
# accuracy of the two hi/lo split variants against float64 (tools/diag_tc.py prints "h2 err vs f64")
echo "== rna split (shipping) =="; python tools/diag_tc.py 2>&1 | grep -A1 "impl 2" | grep "h2 err"
cp super_sac_b200/libssac_b200.so /tmp/keep.so; cp super_sac_b200/libssac_b200_trunc.so super_sac_b200/libssac_b200.so
echo "== truncation split, raw fp32 as hi =="; python tools/diag_tc.py 2>&1 | grep -A8 "impl 2" | grep -E "h2 err|W2|dx"
cp /tmp/keep.so super_sac_b200/libssac_b200.so
