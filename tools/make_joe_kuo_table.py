#!/usr/bin/env python
"""Pack the Joe-Kuo (new-joe-kuo-6.21201) primitive polynomials and initial direction
numbers into a compact binary table used by the host side of salib_b200.

Source of the numbers: the reference's data table /root/reference/src/SALib/sample/directions.py
(tuples ``(a, m_1..m_s)`` for dimensions 2..21201).  Only the *numbers* are taken; the
recurrence that turns them into direction vectors is implemented in
salib_b200/csrc/host_tables.cpp and restated independently in oracle/sobol_seq.py.

Run in the build container only (needs /root/reference):
    python tools/make_joe_kuo_table.py
Writes salib_b200/data/joe_kuo_6_21201.npz with
    a : uint32 [21200]      polynomial coefficient bits (leading/trailing 1 omitted)
    s : uint8  [21200]      polynomial degree
    m : uint32 [21200, 18]  initial direction numbers, zero padded
"""
import os
import sys

import numpy as np

sys.path.insert(0, "/root/reference/src")
from SALib.sample.directions import directions  # noqa: E402

n = len(directions)
a = np.zeros(n, dtype=np.uint32)
s = np.zeros(n, dtype=np.uint8)
m = np.zeros((n, 18), dtype=np.uint32)
for i, tup in enumerate(directions):
    a[i] = tup[0]
    s[i] = len(tup) - 1
    m[i, : len(tup) - 1] = tup[1:]
out = os.path.join(os.path.dirname(__file__), "..", "salib_b200", "data", "joe_kuo_6_21201.npz")
np.savez_compressed(out, a=a, s=s, m=m)
print(n, "dims packed ->", os.path.abspath(out), os.path.getsize(out), "bytes")
