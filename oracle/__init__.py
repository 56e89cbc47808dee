"""CPU oracles for the MADRL env hot paths -- TEST INFRASTRUCTURE ONLY.

Restatements (float64 / integer NumPy) of the reference's algorithms, pinned against the real
reference classes (``refshim.py``) and against the golden vectors in ``tests/golden``.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this
package; the product package ``madrl_b200`` never does.
"""
