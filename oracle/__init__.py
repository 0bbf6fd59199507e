"""oracle/ -- TEST INFRASTRUCTURE ONLY (CPU restatements of the reference algorithms).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package; the product (canonicalvoting_b200, hv_cuda,
hough_voting, MinkowskiEngine shims) never does.
"""
