"""CPU oracle for the PPT tokenizer hot path -- TEST INFRASTRUCTURE ONLY.

May be imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs, and by nothing under ppt_b200/.  Parity status: PINNED to fixtures
generated from the unmodified reference (oracle/gen_golden.py).
"""
