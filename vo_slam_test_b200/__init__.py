"""vo_slam_test_b200 -- B200-native (sm_100a) ORB front-end for guisongchen/vo_slam_test.

Python host mirror of the reference's operator API for the hot path:

* ``ORBextractor``  <->  ORB_SLAM2::ORBextractor  (include/myslam/ORBextractor.h:45-108)
* ``Matcher``       <->  myslam::Matcher          (include/myslam/matcher.h:9-45) -- computeDistance,
                         the best/second-best+ratio loop, searchByProjection(Frame,Frame) and
                         searchByProjection(Frame, local map points)
* ``grid_build``    <->  Frame::assignFeaturesToGrid (src/frame.cpp:72-89)

Everything computes in hand-written CUDA behind the C ABI of ``include/orb_b200.h``
(``vo_slam_test_b200/lib/libvoslam_b200.so``).  There is no CPU fallback: importing works anywhere, but
any compute call without the built library or without a CUDA device raises.
"""
from .api import (KP_DTYPE, Frame, HostBuffer, UploadedFrame, Matcher, ORBextractor, OrbError, camera, device_count, frame_finish, grid_build, knn2_device, knn2_merge_device,
                  knn2_pairs_device, knn2_workspace_bytes, lib, lib_path, load_library, medoid_descriptors)

__all__ = ["ORBextractor", "Matcher", "grid_build", "knn2_device", "KP_DTYPE", "OrbError", "device_count", "lib",
           "lib_path", "load_library", "medoid_descriptors", "camera", "frame_finish"]
