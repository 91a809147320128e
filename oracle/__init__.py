"""Test oracle (CPU restatement of the reference pivot loop). Test infrastructure only:
imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs - never by
network_flow_solver_b200/."""
