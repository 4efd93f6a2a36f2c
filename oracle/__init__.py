"""CPU oracle for the hyperreel_b200 hot path -- TEST INFRASTRUCTURE ONLY.

``hyperreel_oracle.py`` restates the reference's per-ray render algorithm in plain torch CPU arithmetic (pinned to the
unmodified reference through ``ref_shim.py`` and the golden vectors under ``tests/golden/``); ``rays_oracle.py`` restates the
camera -> rays step.  Only ``tests/``, ``__graft_entry__.smoke()`` and the CPU-baseline legs of ``bench.py`` may import this
package; the product (``hyperreel_b200/``) never does.
"""
