#!/bin/bash
# ORACLE / TEST INFRASTRUCTURE: builds the CPU restatement into oracle/_build/libgecut_oracle.so
# (-ffp-contract=off: Julia never fuses a*b+c).  The reference itself is Julia and cannot be compiled here,
# so there is no oracle/_ref.
set -e
cd "$(dirname "$0")"
mkdir -p _build
python emit_tables_inc.py
g++ -O2 -ffp-contract=off -fno-fast-math -std=c++17 -shared -fPIC -o _build/libgecut_oracle.so gecut_oracle.cpp
echo "built oracle/_build/libgecut_oracle.so"
