"""oracle/ -- TEST INFRASTRUCTURE, not product code.

CPU restatements of the rl-agents planning hot path (OPD / MCTS / OLOP /
value iteration), the frozen CPU env models they are driven on, and the
loader that imports the *unmodified* reference from /root/reference in the
build container to pin those restatements.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
arm may import anything from this package, and only as the checker or the
timed CPU baseline.  The product package (rl_agents_b200) never imports it.
"""
