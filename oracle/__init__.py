"""oracle/ -- TEST INFRASTRUCTURE, not product code.

CPU restatement of the reference's coupling-flow algorithm (flow_oracle.py) and the
deterministic synthetic weights/inputs (synth.py) shared by the golden-fixture generator,
the tests and bench.py.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs may import this package; the shipped package
(normalising-flows-4-reinforcement-learning_b200/, imported as ``nf_b200``) never does.
"""
