"""TEST INFRASTRUCTURE ONLY.

CPU restatement ("oracle") of the HM-ViT fusion hot path.  Nothing in the
product package may import this; only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs use it, and only as the checker
or as the timed CPU baseline.
"""
