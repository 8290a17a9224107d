// Host-only: PointCloud3f PLY passthrough (utilities/point_cloud.hpp:118-121, :502-543). No GPU needed.
#include <cilantro/utilities/ply_io.hpp>
#include <cilantro/utilities/point_cloud.hpp>

#include <cstdio>
#include <fstream>

#define CHECK(cond)                                                        \
  do {                                                                     \
    if (!(cond)) {                                                         \
      std::printf("CHECK failed at %s:%d: %s\n", __FILE__, __LINE__, #cond); \
      return 1;                                                            \
    }                                                                      \
  } while (0)

int main(int argc, char** argv) {
  const std::string dir = argc > 1 ? argv[1] : ".";
  if (argc > 2) {
    // a real scan (the reference's examples/test_clouds/*.ply): print what was read for the Python side to check
    cilantro::PointCloud3f scan(argv[2]);
    std::printf("scan %zu %d %d\n", scan.size(), (int)scan.hasNormals(), (int)scan.hasColors());
    const size_t probe[3] = {0, scan.size() / 2, scan.size() - 1};
    for (size_t i : probe) {
      std::printf("v %zu %.9g %.9g %.9g", i, scan.points(0, i), scan.points(1, i), scan.points(2, i));
      if (scan.hasNormals()) std::printf(" n %.9g %.9g %.9g", scan.normals(0, i), scan.normals(1, i), scan.normals(2, i));
      if (scan.hasColors()) std::printf(" c %.9g %.9g %.9g", scan.colors(0, i), scan.colors(1, i), scan.colors(2, i));
      std::printf("\n");
    }
    return 0;
  }
  cilantro::PointCloud3f pc;
  const size_t N = 1000;
  pc.points.resize(3, N);
  pc.normals.resize(3, N);
  pc.colors.resize(3, N);
  for (size_t i = 0; i < N; i++) {
    pc.points.setCol(i, {0.001f * i, -1.5f + 0.25f * (i % 7), 1e-3f * (float)(i * i % 101)});
    pc.normals.setCol(i, {0.f, 0.6f, -0.8f});
    pc.colors.setCol(i, {(i % 256) / 255.0f, ((3 * i) % 256) / 255.0f, 1.0f});
  }
  for (int binary = 0; binary < 2; binary++) {
    const std::string f = dir + (binary ? "/cloud_bin.ply" : "/cloud_ascii.ply");
    pc.toPLYFile(f, binary != 0);
    cilantro::PointCloud3f back(f);
    CHECK(back.size() == N && back.hasNormals() && back.hasColors());
    for (size_t i = 0; i < N; i++)
      for (int r = 0; r < 3; r++) {
        CHECK(back.points(r, i) == pc.points(r, i));    // 9 significant digits round-trip fp32 exactly
        CHECK(back.normals(r, i) == pc.normals(r, i));
        CHECK(std::fabs(back.colors(r, i) - pc.colors(r, i)) <= 1.0f / 255.0f);  // stored as uchar, truncated
      }
  }
  // points only; then a hand-written file with another layout: doubles, extra properties, a face element
  cilantro::PointCloud3f bare;
  bare.points = pc.points;
  bare.toPLYFile(dir + "/bare.ply");
  cilantro::PointCloud3f bare_back(dir + "/bare.ply");
  CHECK(bare_back.size() == N && !bare_back.hasNormals() && !bare_back.hasColors());
  {
    std::ofstream o(dir + "/hand.ply");
    o << "ply\nformat ascii 1.0\ncomment made by hand\nelement vertex 3\nproperty double x\nproperty double y\n"
         "property double z\nproperty uchar red\nproperty uchar green\nproperty uchar blue\nproperty float quality\n"
         "element face 1\nproperty list uchar int vertex_indices\nend_header\n"
         "0 0 0 255 0 0 0.5\n1 0 0 0 255 0 0.25\n0 1 0.5 0 0 51 1\n3 0 1 2\n";
  }
  cilantro::PointCloud3f hand(dir + "/hand.ply");
  CHECK(hand.size() == 3 && hand.hasColors() && !hand.hasNormals());
  CHECK(hand.points(2, 2) == 0.5f && hand.points(0, 1) == 1.f);
  CHECK(hand.colors(0, 0) == 1.f && std::fabs(hand.colors(2, 2) - 0.2f) < 1e-6f);
  // append / clear
  hand.append(hand);
  CHECK(hand.size() == 6 && hand.hasColors());
  hand.clear();
  CHECK(hand.isEmpty());
  bool threw = false;
  try {
    cilantro::PointCloud3f missing(dir + "/does_not_exist.ply");
  } catch (const std::runtime_error&) {
    threw = true;
  }
  CHECK(threw);
  std::printf("PLY checks passed\n");
  return 0;
}
