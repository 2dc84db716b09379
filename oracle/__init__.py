"""
oracle/ — CPU restatement of the reference's gridding path.  TEST INFRASTRUCTURE ONLY.

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs; the product package never imports it (tests/test_no_oracle_in_product.py enforces that).
"""
