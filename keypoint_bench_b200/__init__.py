"""B200-native (sm_100a) post-network hot path of keypoint_bench.

Only the path named in BASELINE.json is here: score-map NMS / border / threshold /
top-k, bilinear descriptor sampling, brute-force mutual-NN matching, homography
projection and the repeatability / MHA counting kernels.  See DESIGN.md.
"""
__version__ = "0.1.0"
