"""CPU oracle for the keypoint_bench post-network hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``keypoint_bench_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs use it, and there only as the checker
(or as the timed CPU baseline), never as the product path.

Parity pinning: the reference ships no tests or golden vectors,
so the restatements in ``oracle/ref_ops.py`` are pinned against OUTPUTS OF THE
REFERENCE ITSELF, imported read-only from /root/reference in the authoring
container by ``oracle/make_golden.py`` (fixtures in ``tests/golden/*.npz``; the same run
asserts live equality of every restatement with the reference and logs it in
``oracle/REFCHECK.log``).  The one third-party function on the path that is absent
from the tree, ``skimage.feature.match_descriptors`` (scikit-image, unpinned in
requirements.txt:13), is restated from its published algorithm over
``scipy.spatial.distance.cdist`` -- that single function is "parity unpinned" by
the reference; every other function is pinned by the committed fixtures.
"""
