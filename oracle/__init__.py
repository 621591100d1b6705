"""CPU oracle for the MindTheEdge depth-edge hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker or the timed
CPU baseline.  ``mindtheedge_b200`` never imports this package and has no CPU
fallback.

Each module restates one part of the reference algorithm and cites the
reference ``file:line`` it follows (paths relative to the reference root).
Pinning status (see DESIGN.md "Oracle"):

* edge_loss      pinned  - golden vectors generated from the reference GradLoss
* canny          pinned  - cv2.Canny itself (the reference's dependency) + goldens
* dee            pinned  - golden vectors generated from the reference tools.py
* pr_counts      PARITY UNPINNED for the matcher: the reference delegates to
                 py-bsds500 ``correspond_pixels`` (not vendored, no version pin,
                 absent here); restated as maximum-cardinality matching and
                 cross-checked with scipy's Hopcroft-Karp.
* thin           PARITY UNPINNED - py-bsds500 ``binary_thin`` restated from its
                 published algorithm (MATLAB bwmorph 'thin' LUTs).
"""
