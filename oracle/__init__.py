"""Oracle: CPU restatements of the reference's hot path (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package.  The product package (recbole-fairrec_b200/) never does.
"""
