// ROS-free replay of the reference node's per-scan callback, written against the reference's own API
// (class NDTFrame, PSOConfig; include <ndtpso_slam/ndtframe.h>, <ndtpso_slam/core.h>): the body of
// NDTPSONode::scan_matcher_ (src/ndtpso_slam_node.cpp:177-244) with the ROS message replaced by a
// binary file of scans.  Compiled against THIS repository's drop-in headers and libndtpso_slam.so it
// runs the matching on the GPU; compiled against the reference's headers and library the same file
// runs on the CPU (only the 3-argument align() overload below is an addition of the drop-in).
//
//   replay_node <scans.bin> [iterations population]
//   scans.bin: int32 n_scans, int32 n_beams, float angle_min, angle_increment, range_max,
//              double initial_pose[3], int32 frame_size_m, double cell_side, then n_scans x n_beams float ranges
//   stdout:    one "x y theta" line per scan (%.17g)
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <ndtpso_slam/core.h>
#include <ndtpso_slam/ndtframe.h>

int main(int argc, char** argv) {
  if (argc < 2) {
    fprintf(stderr, "usage: %s scans.bin [iterations population]\n", argv[0]);
    return 2;
  }
  FILE* f = fopen(argv[1], "rb");
  if (!f) {
    perror(argv[1]);
    return 2;
  }
  int32_t n_scans = 0, n_beams = 0, frame_size = 0;
  float angle_min = 0, angle_increment = 0, range_max = 0;
  double init[3], cell_side = 0;
  bool ok = fread(&n_scans, 4, 1, f) == 1 && fread(&n_beams, 4, 1, f) == 1 && fread(&angle_min, 4, 1, f) == 1 &&
            fread(&angle_increment, 4, 1, f) == 1 && fread(&range_max, 4, 1, f) == 1 && fread(init, 8, 3, f) == 3 &&
            fread(&frame_size, 4, 1, f) == 1 && fread(&cell_side, 8, 1, f) == 1;
  std::vector<std::vector<float>> scans(ok ? n_scans : 0, std::vector<float>(ok ? n_beams : 0));
  for (auto& s : scans) ok = ok && fread(s.data(), 4, s.size(), f) == s.size();
  fclose(f);
  if (!ok) {
    fprintf(stderr, "%s: truncated\n", argv[1]);
    return 2;
  }
  PSOConfig pso;  // the node fills this from its ROS parameters (ndtpso_slam_node.cpp:30-36)
  if (argc >= 4) {
    pso.iterations = atoi(argv[2]);
    pso.populationSize = atoi(argv[3]);
  }

  // ---- NDTPSONode::NDTPSONode (ndtpso_slam_node.cpp:64-78)
  const Vector3d initial_pose(init[0], init[1], init[2]);
  const unsigned short size = static_cast<unsigned short>(frame_size);
  NDTFrame* ref_frame = new NDTFrame(Vector3d::Zero(), size, size, cell_side, true);
  NDTFrame* current_frame = new NDTFrame(initial_pose, size, size, cell_side, false);
  Vector3d previous_pose = initial_pose, current_pose = initial_pose;
  bool first_iteration = true;

  // ---- NDTPSONode::scan_matcher_ (ndtpso_slam_node.cpp:177-244), once per scan
  for (const auto& ranges : scans) {
    current_frame->loadLaser(ranges, angle_min, angle_increment, range_max);
    if (first_iteration)
      current_pose = previous_pose;
    else
      current_pose = (argc >= 4) ? ref_frame->align(previous_pose, current_frame, pso)  // the overload that honours the parameters
                                 : ref_frame->align(previous_pose, current_frame);      // the reference's call (:194)
    previous_pose = current_pose;
    ref_frame->update(current_pose, current_frame);
    printf("%.17g %.17g %.17g\n", current_pose.x(), current_pose.y(), current_pose.z());
    delete current_frame;
    current_frame = new NDTFrame(initial_pose, size, size, static_cast<double>(size), false);
    first_iteration = false;
  }
  delete current_frame;
  delete ref_frame;
  return 0;
}
