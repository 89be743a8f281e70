"""vieo_slam_b200 — B200-native hot path of VIEO_SLAM (ORB front-end, Hamming matching, IMU
pre-integration, PoseOptimization / BA) behind a C ABI (include/vieo_b200.h)."""
__version__ = "0.1.0"
