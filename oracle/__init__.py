"""ORACLE package (test infrastructure, NOT product code).

CPU restatement of the reference's self-play hot path; see omok_oracle.h,
omok_oracle.c, net_oracle.py.  Importable only from tests/, smoke() and
bench.py's cpu_baseline / --impl reference legs.
"""
