"""TEST INFRASTRUCTURE ONLY.

CPU restatement (torch fp32) of the reference hot path plus a shim that imports the
real reference when /root/reference is present.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this package; the
product package (vae_segmentation_b200) never does.
"""
