"""CPU oracle for the Gnet hot path: TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs import this package; gossipnet_b200 never does.
"""
