"""CPU oracle for the quill-zkvm hot path (TEST INFRASTRUCTURE, not product code).

PARITY UNPINNED: see oracle/README.md.  Import only from tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs.
"""
