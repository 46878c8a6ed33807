"""ORACLE — test infrastructure, not product code.

CPU restatement of the reference hot path (VladGavra98/SERL: Agent.evaluate x population -> fitness ->
SSNE.epoch) used ONLY as the checker in tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs.  Nothing under serl_b200/ may import this package.

Pinning (see DESIGN.md "Oracle"):
  * plant      : oracle/_ref/citation_<variant>.so are byte copies of the reference's own native plant
                 binaries (made by oracle/build.py when /root/reference is present); the C restatement
                 oracle/plant/plant_oracle.c is bit-identical to them on the 15 logged reference episodes
                 and on random right-hand-side evaluations (tests/golden/plant_rhs_kat.npz).
  * actor      : restated from base/core/genetic_agent.py:69-109, pinned by the logged TD3 episode 575.
  * env wrapper: restated from envs/phlabenv.py:62-73, 347-486; reward column of the logged episodes.
  * reference-signal generator: the third-party `signals==0.0.1` package is absent -> PARITY UNPINNED for
                 its RNG stream; oracle and product are fed identical, explicitly parameterised signals.
  * EA         : base/core/mod_neuro_evo.py restated with the exclusive-index patch, checked
                 against the reference module itself (imported from /root/reference in the container).
"""
