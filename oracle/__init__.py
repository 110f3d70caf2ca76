"""CPU oracle for the two ii-vision hot paths.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this package; the product (iivision_b200) never does.
"""
