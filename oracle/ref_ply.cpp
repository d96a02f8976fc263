// TEST INFRASTRUCTURE ONLY.  PLY oracle for row f-3's on-disk format: drives the reference's OWN vendored
// PLY library (submodules/utils/ply_utils/tinyply.h, compiled from where it lies under /root/reference)
// with exactly the property sequence of GaussianModel::savePly (src/gaussian_model.cpp:861-953) and the
// property requests of GaussianModel::loadPly (:688-859).  Built by oracle/build_ref_ply.sh into
// oracle/_ref/ply_ref; tests/golden/make_golden_ply.py used it to write tests/golden/gaussians_*.ply.
//
//   ply_ref write <raw.bin> <out.ply>   raw.bin: int32 P, int32 M, then the tensors in MODEL layout:
//                                        xyz[P,3] f_dc[P,1,3] f_rest[P,M,3] opacity[P,1] scale[P,3] rot[P,4]
//   ply_ref read  <in.ply> <raw.bin> <sh_degree>   the inverse (loadPly's transposes applied)
#define TINYPLY_IMPLEMENTATION
#include "tinyply.h"

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

static std::vector<std::string> names(const char* prefix, int n) {
  std::vector<std::string> v(n);
  for (int i = 0; i < n; ++i) v[i] = std::string(prefix) + std::to_string(i);
  return v;
}

// tensor.transpose(1, 2).flatten(1) of a [P, A, B] array -> [P, B*A]
static std::vector<float> transpose12(const std::vector<float>& t, int P, int A, int B) {
  std::vector<float> o(t.size());
  for (int p = 0; p < P; ++p)
    for (int a = 0; a < A; ++a)
      for (int b = 0; b < B; ++b) o[(size_t)p * A * B + (size_t)b * A + a] = t[(size_t)p * A * B + (size_t)a * B + b];
  return o;
}

static int do_write(const char* raw, const char* out) {
  std::ifstream in(raw, std::ios::binary);
  int32_t P = 0, M = 0;
  in.read(reinterpret_cast<char*>(&P), 4);
  in.read(reinterpret_cast<char*>(&M), 4);
  auto rd = [&](size_t n) { std::vector<float> v(n); in.read(reinterpret_cast<char*>(v.data()), n * 4); return v; };
  std::vector<float> xyz = rd((size_t)P * 3), dc = rd((size_t)P * 3), rest = rd((size_t)P * M * 3),
                     opac = rd(P), scale = rd((size_t)P * 3), rot = rd((size_t)P * 4);
  std::vector<float> normals((size_t)P * 3, 0.f);                 // :866
  std::vector<float> f_dc = transpose12(dc, P, 1, 3);             // :867
  std::vector<float> f_rest = transpose12(rest, P, M, 3);         // :868
  std::filebuf fb;
  fb.open(out, std::ios::out | std::ios::binary);
  std::ostream os(&fb);
  tinyply::PlyFile f;
  auto add = [&](const std::vector<std::string>& n, std::vector<float>& d) {
    f.add_properties_to_element("vertex", n, tinyply::Type::FLOAT32, P, reinterpret_cast<uint8_t*>(d.data()),
                                tinyply::Type::INVALID, 0);
  };
  add({"x", "y", "z"}, xyz);              // :881-885
  add({"nx", "ny", "nz"}, normals);       // :888-892
  add(names("f_dc_", 3), f_dc);           // :895-904
  add(names("f_rest_", 3 * M), f_rest);   // :907-916
  add({"opacity"}, opac);                 // :919-923
  add(names("scale_", 3), scale);         // :926-935
  add(names("rot_", 4), rot);             // :938-947
  f.write(os, true);                      // :950
  fb.close();
  return 0;
}

static int do_read(const char* ply, const char* raw, int degree) {
  std::ifstream is(ply, std::ios::binary);
  tinyply::PlyFile f;
  f.parse_header(is);
  const int n_rest = ((degree + 1) * (degree + 1) - 1) * 3;       // :738
  auto xyz = f.request_properties_from_element("vertex", {"x", "y", "z"});
  auto dc = f.request_properties_from_element("vertex", {"f_dc_0", "f_dc_1", "f_dc_2"});
  auto rest = f.request_properties_from_element("vertex", names("f_rest_", n_rest));
  auto opac = f.request_properties_from_element("vertex", {"opacity"});
  auto scale = f.request_properties_from_element("vertex", {"scale_0", "scale_1", "scale_2"});
  auto rot = f.request_properties_from_element("vertex", {"rot_0", "rot_1", "rot_2", "rot_3"});
  f.read(is);
  const int32_t P = (int32_t)xyz->count, M = n_rest / 3;
  auto vec = [&](std::shared_ptr<tinyply::PlyData>& d, size_t n) {
    std::vector<float> v(n);
    std::memcpy(v.data(), d->buffer.get(), n * 4);
    return v;
  };
  // from_blob(.., {P, 3, k}).transpose(1, 2).contiguous() (:829-843): [P,3,k] -> [P,k,3]
  std::vector<float> dcv = transpose12(vec(dc, (size_t)P * 3), P, 3, 1);
  std::vector<float> restv = transpose12(vec(rest, (size_t)P * n_rest), P, 3, M);
  std::ofstream o(raw, std::ios::binary);
  o.write(reinterpret_cast<const char*>(&P), 4);
  o.write(reinterpret_cast<const char*>(&M), 4);
  auto wr = [&](const std::vector<float>& v) { o.write(reinterpret_cast<const char*>(v.data()), v.size() * 4); };
  wr(vec(xyz, (size_t)P * 3)); wr(dcv); wr(restv); wr(vec(opac, P)); wr(vec(scale, (size_t)P * 3)); wr(vec(rot, (size_t)P * 4));
  return 0;
}

int main(int argc, char** argv) {
  if (argc >= 4 && std::string(argv[1]) == "write") return do_write(argv[2], argv[3]);
  if (argc >= 5 && std::string(argv[1]) == "read") return do_read(argv[2], argv[3], std::atoi(argv[4]));
  std::fprintf(stderr, "usage: ply_ref write raw.bin out.ply | ply_ref read in.ply raw.bin sh_degree\n");
  return 2;
}
