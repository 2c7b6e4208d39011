// scene_loader.cpp -- the reference's text scene format, parsed with the reference's semantics.
//
// Follows (paths relative to the reference repo root):
//   scene::scene / loadMaterial / loadObject / loadCamera   src/scene.cpp:11-35, 222-265, 37-135, 137-220
//   utilityCore::safeGetline / tokenizeString               src/utilities.cpp:109-139, 101-107
//   utilityCore::buildTransformationMatrix                  src/utilities.cpp:74-81   (T * Rx * Ry * Rz * S)
//   utilityCore::glmMat4ToCudaMat4                          src/utilities.cpp:83-90   (transpose -> 4 rows)
//   GLM 0.9.5.4 translate / rotate / scale / operator* / inverse, whose float operation order is reproduced so
//   that matrices come out bit-identical to the reference loader's:
//       external/include/glm/gtc/matrix_transform.inl:35-90,128-141
//       external/include/glm/detail/type_mat4x4.inl:476-531 (inverse), 753-775 (operator*)
//
// Reference quirks kept on purpose:
//   * ROTAT is consumed as RADIANS (utilities.cpp:7 defines GLM_FORCE_RADIANS)          -- SURVEY.md D1
//     (rotat_degrees=1 converts first, for scenes authored in degrees)
//   * the object type line must equal "sphere" / "cube" exactly (scene.cpp:50-55)
//   * a MATERIAL block is exactly 10 lines, a CAMERA block exactly 4 static lines (scene.cpp:143,232)
//   * unknown top-level lines are ignored (scene.cpp:22-31) -- that is where the new LENS block lives
// Differences, on purpose: errors are returned (PT_ERR_PARSE + message) instead of printed and ignored; values a
// block does not mention are 0 instead of uninitialised; nothing is printed.
#include "../../include/pt_b200.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

extern "C" void pt_set_error_(const char* fmt, ...);

namespace {

struct V3 { float x, y, z; };
struct V4 { float x, y, z, w; };
struct M4 { V4 c[4]; };  // column-major like glm::mat4

inline V4 mulv(V4 a, float s) { return V4{a.x * s, a.y * s, a.z * s, a.w * s}; }
inline V4 addv(V4 a, V4 b) { return V4{a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }
inline V4 subv(V4 a, V4 b) { return V4{a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w}; }
inline V4 mulvv(V4 a, V4 b) { return V4{a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w}; }
inline float at(const V4& v, int i) { return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w; }

M4 identity() {
  M4 m;
  m.c[0] = V4{1, 0, 0, 0}; m.c[1] = V4{0, 1, 0, 0}; m.c[2] = V4{0, 0, 1, 0}; m.c[3] = V4{0, 0, 0, 1};
  return m;
}

// glm::translate(m, v): Result[3] = m[0]*v[0] + m[1]*v[1] + m[2]*v[2] + m[3]
M4 translate(const M4& m, V3 v) {
  M4 r = m;
  r.c[3] = addv(addv(addv(mulv(m.c[0], v.x), mulv(m.c[1], v.y)), mulv(m.c[2], v.z)), m.c[3]);
  return r;
}

// glm::rotate(m, angle, axis) with GLM_FORCE_RADIANS
M4 rotate(const M4& m, float angle, V3 v) {
  const float a = angle;
  const float c = std::cos(a);
  const float s = std::sin(a);
  const float sqr = v.x * v.x + v.y * v.y + v.z * v.z;
  const float inv = 1.0f / std::sqrt(sqr);
  const float axis[3] = {v.x * inv, v.y * inv, v.z * inv};
  const float temp[3] = {(1.0f - c) * axis[0], (1.0f - c) * axis[1], (1.0f - c) * axis[2]};
  float R[3][3];
  R[0][0] = c + temp[0] * axis[0];
  R[0][1] = 0 + temp[0] * axis[1] + s * axis[2];
  R[0][2] = 0 + temp[0] * axis[2] - s * axis[1];
  R[1][0] = 0 + temp[1] * axis[0] - s * axis[2];
  R[1][1] = c + temp[1] * axis[1];
  R[1][2] = 0 + temp[1] * axis[2] + s * axis[0];
  R[2][0] = 0 + temp[2] * axis[0] + s * axis[1];
  R[2][1] = 0 + temp[2] * axis[1] - s * axis[0];
  R[2][2] = c + temp[2] * axis[2];
  M4 r;
  for (int j = 0; j < 3; j++)
    r.c[j] = addv(addv(mulv(m.c[0], R[j][0]), mulv(m.c[1], R[j][1])), mulv(m.c[2], R[j][2]));
  r.c[3] = m.c[3];
  return r;
}

// glm::scale(m, v)
M4 scale(const M4& m, V3 v) {
  M4 r;
  r.c[0] = mulv(m.c[0], v.x); r.c[1] = mulv(m.c[1], v.y); r.c[2] = mulv(m.c[2], v.z); r.c[3] = m.c[3];
  return r;
}

// tmat4x4 operator*: Result[j] = A0*B[j][0] + A1*B[j][1] + A2*B[j][2] + A3*B[j][3]
M4 mul(const M4& A, const M4& B) {
  M4 r;
  for (int j = 0; j < 4; j++)
    r.c[j] = addv(addv(addv(mulv(A.c[0], B.c[j].x), mulv(A.c[1], B.c[j].y)), mulv(A.c[2], B.c[j].z)),
                  mulv(A.c[3], B.c[j].w));
  return r;
}

// glm::inverse(mat4): cofactor expansion, compute_inverse<tmat4x4>
M4 inverse(const M4& mm) {
  float m[4][4];
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) m[i][j] = at(mm.c[i], j);
  float Coef00 = m[2][2] * m[3][3] - m[3][2] * m[2][3];
  float Coef02 = m[1][2] * m[3][3] - m[3][2] * m[1][3];
  float Coef03 = m[1][2] * m[2][3] - m[2][2] * m[1][3];
  float Coef04 = m[2][1] * m[3][3] - m[3][1] * m[2][3];
  float Coef06 = m[1][1] * m[3][3] - m[3][1] * m[1][3];
  float Coef07 = m[1][1] * m[2][3] - m[2][1] * m[1][3];
  float Coef08 = m[2][1] * m[3][2] - m[3][1] * m[2][2];
  float Coef10 = m[1][1] * m[3][2] - m[3][1] * m[1][2];
  float Coef11 = m[1][1] * m[2][2] - m[2][1] * m[1][2];
  float Coef12 = m[2][0] * m[3][3] - m[3][0] * m[2][3];
  float Coef14 = m[1][0] * m[3][3] - m[3][0] * m[1][3];
  float Coef15 = m[1][0] * m[2][3] - m[2][0] * m[1][3];
  float Coef16 = m[2][0] * m[3][2] - m[3][0] * m[2][2];
  float Coef18 = m[1][0] * m[3][2] - m[3][0] * m[1][2];
  float Coef19 = m[1][0] * m[2][2] - m[2][0] * m[1][2];
  float Coef20 = m[2][0] * m[3][1] - m[3][0] * m[2][1];
  float Coef22 = m[1][0] * m[3][1] - m[3][0] * m[1][1];
  float Coef23 = m[1][0] * m[2][1] - m[2][0] * m[1][1];
  V4 Fac0{Coef00, Coef00, Coef02, Coef03}, Fac1{Coef04, Coef04, Coef06, Coef07}, Fac2{Coef08, Coef08, Coef10, Coef11};
  V4 Fac3{Coef12, Coef12, Coef14, Coef15}, Fac4{Coef16, Coef16, Coef18, Coef19}, Fac5{Coef20, Coef20, Coef22, Coef23};
  V4 Vec0{m[1][0], m[0][0], m[0][0], m[0][0]}, Vec1{m[1][1], m[0][1], m[0][1], m[0][1]};
  V4 Vec2{m[1][2], m[0][2], m[0][2], m[0][2]}, Vec3{m[1][3], m[0][3], m[0][3], m[0][3]};
  V4 Inv0 = addv(subv(mulvv(Vec1, Fac0), mulvv(Vec2, Fac1)), mulvv(Vec3, Fac2));
  V4 Inv1 = addv(subv(mulvv(Vec0, Fac0), mulvv(Vec2, Fac3)), mulvv(Vec3, Fac4));
  V4 Inv2 = addv(subv(mulvv(Vec0, Fac1), mulvv(Vec1, Fac3)), mulvv(Vec3, Fac5));
  V4 Inv3 = addv(subv(mulvv(Vec0, Fac2), mulvv(Vec1, Fac4)), mulvv(Vec2, Fac5));
  V4 SignA{+1, -1, +1, -1}, SignB{-1, +1, -1, +1};
  M4 Inverse;
  Inverse.c[0] = mulvv(Inv0, SignA); Inverse.c[1] = mulvv(Inv1, SignB);
  Inverse.c[2] = mulvv(Inv2, SignA); Inverse.c[3] = mulvv(Inv3, SignB);
  V4 Row0{Inverse.c[0].x, Inverse.c[1].x, Inverse.c[2].x, Inverse.c[3].x};
  V4 Dot0 = mulvv(mm.c[0], Row0);
  float Dot1 = (Dot0.x + Dot0.y) + (Dot0.z + Dot0.w);
  float OneOverDeterminant = 1.0f / Dot1;
  M4 r;
  for (int j = 0; j < 4; j++) r.c[j] = mulv(Inverse.c[j], OneOverDeterminant);
  return r;
}

// glmMat4ToCudaMat4: rows of the matrix, i.e. the transpose of glm's column storage
void to_rows(const M4& a, float out[16]) {
  for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) out[4 * r + c] = at(a.c[c], r);
}

M4 build_transform(V3 t, V3 r, V3 s) {
  M4 translationMat = translate(identity(), t);
  M4 rotationMat = rotate(identity(), r.x, V3{1, 0, 0});
  rotationMat = mul(rotationMat, rotate(identity(), r.y, V3{0, 1, 0}));
  rotationMat = mul(rotationMat, rotate(identity(), r.z, V3{0, 0, 1}));
  M4 scaleMat = scale(identity(), s);
  return mul(mul(translationMat, rotationMat), scaleMat);
}

struct Object {
  int type = 0, materialid = 0;
  std::vector<V3> translations, rotations, scales;
  std::vector<M4> transforms, inverses;
};

struct Reader {
  std::string data;
  size_t pos = 0;
  bool eof = false;
  // safeGetline: a line ends at \n, \r\n, \r or end of file; good() turns false once EOF is hit on an empty line
  bool good() const { return !eof; }
  void getline(std::string& t) {
    t.clear();
    for (;;) {
      if (pos >= data.size()) { if (t.empty()) eof = true; return; }
      char c = data[pos++];
      if (c == '\n') return;
      if (c == '\r') { if (pos < data.size() && data[pos] == '\n') pos++; return; }
      t += c;
    }
  }
};

std::vector<std::string> tokenize(const std::string& s) {
  std::istringstream ss(s);
  std::vector<std::string> out;
  std::string tok;
  while (ss >> tok) out.push_back(tok);
  return out;
}

bool vec3_of(const std::vector<std::string>& tok, V3* v) {
  if (tok.size() < 4) return false;
  *v = V3{(float)atof(tok[1].c_str()), (float)atof(tok[2].c_str()), (float)atof(tok[3].c_str())};
  return true;
}

}  // namespace

struct pt_scene {
  std::vector<Object> objects;
  std::vector<pt_material> materials;
  int width = 0, height = 0, iterations = 0;
  float fov[2] = {0, 0};
  std::string image_name;
  std::vector<V3> eyes, views, ups;
  pt_lens lens{0.0f, 0.0f};
  bool have_camera = false;
};

#define PARSE_FAIL(...)             \
  do {                              \
    pt_set_error_(__VA_ARGS__);     \
    delete s;                       \
    return PT_ERR_PARSE;            \
  } while (0)

extern "C" int pt_scene_load(const char* path, int rotat_degrees, pt_scene** out) {
  if (!path || !out) { pt_set_error_("path/out is NULL"); return PT_ERR_INVALID; }
  *out = nullptr;
  std::ifstream f(path, std::ios::binary);
  if (!f.is_open()) { pt_set_error_("cannot open scene file %s", path); return PT_ERR_IO; }
  std::stringstream buf;
  buf << f.rdbuf();
  Reader rd;
  rd.data = buf.str();
  pt_scene* s = new pt_scene();
  std::string line;
  while (rd.good()) {
    rd.getline(line);
    if (line.empty()) continue;
    std::vector<std::string> tok = tokenize(line);
    if (tok.empty()) continue;
    if (tok[0] == "MATERIAL") {
      // scene.cpp:222-265
      int id = tok.size() > 1 ? atoi(tok[1].c_str()) : -1;
      if (id != (int)s->materials.size()) PARSE_FAIL("MATERIAL id %d does not match expected %zu", id, s->materials.size());
      pt_material m;
      memset(&m, 0, sizeof(m));
      for (int i = 0; i < 10; i++) {
        rd.getline(line);
        std::vector<std::string> t = tokenize(line);
        if (t.empty()) PARSE_FAIL("MATERIAL %d: expected 10 property lines, got %d", id, i);
        V3 v;
        const std::string& k = t[0];
        auto f1 = [&](float* dst) { if (t.size() > 1) *dst = (float)atof(t[1].c_str()); };
        if (k == "RGB") { if (vec3_of(t, &v)) { m.color[0] = v.x; m.color[1] = v.y; m.color[2] = v.z; } }
        else if (k == "SPECEX") f1(&m.specularExponent);
        else if (k == "SPECRGB") { if (vec3_of(t, &v)) { m.specularColor[0] = v.x; m.specularColor[1] = v.y; m.specularColor[2] = v.z; } }
        else if (k == "REFL") f1(&m.hasReflective);
        else if (k == "REFR") f1(&m.hasRefractive);
        else if (k == "REFRIOR") f1(&m.indexOfRefraction);
        else if (k == "SCATTER") f1(&m.hasScatter);
        else if (k == "ABSCOEFF") { if (vec3_of(t, &v)) { m.absorptionCoefficient[0] = v.x; m.absorptionCoefficient[1] = v.y; m.absorptionCoefficient[2] = v.z; } }
        else if (k == "RSCTCOEFF") f1(&m.reducedScatterCoefficient);
        else if (k == "EMITTANCE") f1(&m.emittance);
      }
      s->materials.push_back(m);
    } else if (tok[0] == "OBJECT") {
      // scene.cpp:37-135
      int id = tok.size() > 1 ? atoi(tok[1].c_str()) : -1;
      if (id != (int)s->objects.size()) PARSE_FAIL("OBJECT id %d does not match expected %zu", id, s->objects.size());
      Object o;
      rd.getline(line);
      if (line == "sphere") o.type = 0;
      else if (line == "cube") o.type = 1;
      else {
        std::string name, ext;
        std::istringstream ls(line);
        std::getline(ls, name, '.');
        std::getline(ls, ext, '.');
        if (ext == "obj") o.type = 2;
        else PARSE_FAIL("OBJECT %d: '%s' is not a valid object type", id, line.c_str());
      }
      rd.getline(line);
      {
        std::vector<std::string> t = tokenize(line);
        if (t.size() < 2) PARSE_FAIL("OBJECT %d: expected 'material <id>'", id);
        o.materialid = atoi(t[1].c_str());
      }
      int frameCount = 0;
      rd.getline(line);
      while (!line.empty() && rd.good()) {
        std::vector<std::string> t = tokenize(line);
        if (t.size() < 2 || t[0] != "frame" || atoi(t[1].c_str()) != frameCount)
          PARSE_FAIL("OBJECT %d: incorrect frame count at '%s'", id, line.c_str());
        V3 tr{0, 0, 0}, ro{0, 0, 0}, sc{0, 0, 0};
        for (int i = 0; i < 3; i++) {
          rd.getline(line);
          t = tokenize(line);
          V3 v;
          if (t.empty() || !vec3_of(t, &v)) PARSE_FAIL("OBJECT %d frame %d: expected TRANS/ROTAT/SCALE x y z", id, frameCount);
          if (t[0] == "TRANS") tr = v;
          else if (t[0] == "ROTAT") ro = v;
          else if (t[0] == "SCALE") sc = v;
        }
        o.translations.push_back(tr); o.rotations.push_back(ro); o.scales.push_back(sc);
        frameCount++;
        rd.getline(line);
      }
      if (frameCount == 0) PARSE_FAIL("OBJECT %d has no frames", id);
      for (int i = 0; i < frameCount; i++) {
        V3 r = o.rotations[i];
        if (rotat_degrees) {
          const float k = 0.01745329251994329576923690768489f;  // glm::radians
          r = V3{r.x * k, r.y * k, r.z * k};
        }
        M4 m = build_transform(o.translations[i], r, o.scales[i]);
        o.transforms.push_back(m);
        o.inverses.push_back(inverse(m));
      }
      s->objects.push_back(o);
    } else if (tok[0] == "CAMERA") {
      // scene.cpp:137-220
      float fovy = 0;
      for (int i = 0; i < 4; i++) {
        rd.getline(line);
        std::vector<std::string> t = tokenize(line);
        if (t.empty()) PARSE_FAIL("CAMERA: expected 4 static lines (RES, FOVY, ITERATIONS, FILE)");
        if (t[0] == "RES" && t.size() >= 3) { s->width = atoi(t[1].c_str()); s->height = atoi(t[2].c_str()); }
        else if (t[0] == "FOVY" && t.size() >= 2) fovy = (float)atof(t[1].c_str());
        else if (t[0] == "ITERATIONS" && t.size() >= 2) s->iterations = atoi(t[1].c_str());
        else if (t[0] == "FILE" && t.size() >= 2) s->image_name = t[1];
      }
      int frameCount = 0;
      rd.getline(line);
      s->eyes.clear(); s->views.clear(); s->ups.clear();
      while (!line.empty() && rd.good()) {
        std::vector<std::string> t = tokenize(line);
        if (t.size() < 2 || t[0] != "frame" || atoi(t[1].c_str()) != frameCount)
          PARSE_FAIL("CAMERA: incorrect frame count at '%s'", line.c_str());
        V3 e{0, 0, 0}, vw{0, 0, 0}, up{0, 0, 0};
        for (int i = 0; i < 3; i++) {
          rd.getline(line);
          t = tokenize(line);
          V3 v;
          if (t.empty() || !vec3_of(t, &v)) PARSE_FAIL("CAMERA frame %d: expected EYE/VIEW/UP x y z", frameCount);
          if (t[0] == "EYE") e = v;
          else if (t[0] == "VIEW") vw = v;
          else if (t[0] == "UP") up = v;
        }
        s->eyes.push_back(e); s->views.push_back(vw); s->ups.push_back(up);
        frameCount++;
        rd.getline(line);
      }
      if (s->width <= 0 || s->height <= 0) PARSE_FAIL("CAMERA: bad RES %d %d", s->width, s->height);
      if (frameCount == 0) PARSE_FAIL("CAMERA has no frames");
      // scene.cpp:203-207: fovy is promoted to double for tan, atan stays float, the division by PI is double
      const double PI = 3.1415926535897932384626422832795028841971;
      float yscaled = (float)tan(fovy * (PI / 180));
      float xscaled = (yscaled * (float)s->width) / (float)s->height;
      float fovx = (float)((atanf(xscaled) * 180) / PI);
      s->fov[0] = fovx;
      s->fov[1] = fovy;
      s->have_camera = true;
    } else if (tok[0] == "LENS") {
      // new block (ignored by the reference's dispatcher, scene.cpp:22-31): key/value lines until an empty line
      rd.getline(line);
      while (!line.empty() && rd.good()) {
        std::vector<std::string> t = tokenize(line);
        if (t.size() >= 2 && t[0] == "APERTURE") s->lens.aperture = (float)atof(t[1].c_str());
        else if (t.size() >= 2 && t[0] == "FOCALDIST") s->lens.focal_distance = (float)atof(t[1].c_str());
        rd.getline(line);
      }
    }
  }
  if (!s->have_camera) PARSE_FAIL("scene has no CAMERA block");
  if (s->objects.empty()) PARSE_FAIL("scene has no OBJECT");
  if (s->materials.empty()) PARSE_FAIL("scene has no MATERIAL");
  *out = s;
  return PT_OK;
}

extern "C" int pt_scene_free(pt_scene* s) {
  delete s;
  return PT_OK;
}

extern "C" int pt_scene_info(const pt_scene* s, int* n_geoms, int* n_materials, int* n_frames, int* width,
                             int* height, int* iterations, char* image_name, int image_name_cap) {
  if (!s) { pt_set_error_("scene is NULL"); return PT_ERR_INVALID; }
  if (n_geoms) *n_geoms = (int)s->objects.size();
  if (n_materials) *n_materials = (int)s->materials.size();
  if (n_frames) *n_frames = (int)s->eyes.size();
  if (width) *width = s->width;
  if (height) *height = s->height;
  if (iterations) *iterations = s->iterations;
  if (image_name && image_name_cap > 0) {
    strncpy(image_name, s->image_name.c_str(), image_name_cap - 1);
    image_name[image_name_cap - 1] = 0;
  }
  return PT_OK;
}

// the flattening loop of cudaRaytraceCore, src/raytraceKernel.cu:123-146.  An object with fewer frames than the
// camera keeps its last frame (the reference would read out of bounds).
extern "C" int pt_scene_frame(const pt_scene* s, int frame, pt_static_geom* geoms, pt_material* materials,
                              pt_camera_data* cam, pt_lens* lens) {
  if (!s) { pt_set_error_("scene is NULL"); return PT_ERR_INVALID; }
  if (frame < 0 || frame >= (int)s->eyes.size()) { pt_set_error_("frame %d outside [0,%zu)", frame, s->eyes.size()); return PT_ERR_INVALID; }
  if (geoms) {
    for (size_t i = 0; i < s->objects.size(); i++) {
      const Object& o = s->objects[i];
      const int f = frame < (int)o.transforms.size() ? frame : (int)o.transforms.size() - 1;
      pt_static_geom g;
      memset(&g, 0, sizeof(g));
      g.type = o.type;
      g.materialid = o.materialid;
      g.translation[0] = o.translations[f].x; g.translation[1] = o.translations[f].y; g.translation[2] = o.translations[f].z;
      g.rotation[0] = o.rotations[f].x; g.rotation[1] = o.rotations[f].y; g.rotation[2] = o.rotations[f].z;
      g.scale[0] = o.scales[f].x; g.scale[1] = o.scales[f].y; g.scale[2] = o.scales[f].z;
      to_rows(o.transforms[f], g.transform);
      to_rows(o.inverses[f], g.inverseTransform);
      geoms[i] = g;
    }
  }
  if (materials) memcpy(materials, s->materials.data(), s->materials.size() * sizeof(pt_material));
  if (cam) {
    cam->resolution[0] = (float)s->width; cam->resolution[1] = (float)s->height;
    cam->position[0] = s->eyes[frame].x; cam->position[1] = s->eyes[frame].y; cam->position[2] = s->eyes[frame].z;
    cam->view[0] = s->views[frame].x; cam->view[1] = s->views[frame].y; cam->view[2] = s->views[frame].z;
    cam->up[0] = s->ups[frame].x; cam->up[1] = s->ups[frame].y; cam->up[2] = s->ups[frame].z;
    cam->fov[0] = s->fov[0]; cam->fov[1] = s->fov[1];
  }
  if (lens) *lens = s->lens;
  return PT_OK;
}
