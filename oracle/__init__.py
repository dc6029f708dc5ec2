"""oracle/ -- TEST INFRASTRUCTURE ONLY (CPU restatement + verbatim build of the reference ray query).

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
"""
