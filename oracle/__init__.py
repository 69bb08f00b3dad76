"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the dGPMP2 inner Gauss-Newton step.

Nothing in the product package (``dgpmp2_b200`` / ``diff_gpmp2``) imports this
package.  Only ``tests/``, ``__graft_entry__.smoke()`` and the CPU-baseline /
``--impl reference`` legs of ``bench.py`` may use it, and only as the checker
or as the timed CPU baseline -- never as the product path.
"""
