// scene.cu -- the path-traced scene pass (BASELINE config 5's "pathtraced scene through the camera") on the
// device: what PathTracer::raytrace_pixel (src/pathtracer/pathtracer.cpp:819-899) adds to sampleBuffer BESIDE the ghosts and
// the starburst -- est_radiance_global_illumination (:279-302), which in the reference as it stands is
//     zero_bounce_radiance (:213-218: the hit surface's emission)
//   + one_bounce_radiance  (:220-231 -> estimate_direct_lighting_importance :136-211: every light sampled, shadow ray, f cos / pdf)
// (the indirect bounces are commented out there, :299).  Delta lights (DirectionalLight / PointLight, scene/light.cpp) take
// one sample with pdf 1, so the estimate is deterministic; area lights draw random samples in the reference and are not
// supported here.  Primary rays: Camera::generate_ray (camera.cpp:278-305) through the pixel centres (raytrace_pixel jitters
// them with the process-global RNG).  Geometry: Triangle::intersect (scene/triangle.cpp:26-113, Moeller-Trumbore, interpolated
// unit normals) and Sphere::intersect (scene/sphere.cpp:11-108); BSDFs: DiffuseBSDF (f = reflectance / pi) and EmissionBSDF.
//
// The acceleration structure is ours: a binary BVH built on the host (median split of the centroids along the longest axis,
// leaves of <= 4 primitives, children adjacent), walked on the device with a per-thread stack; the reference's BVHAccel
// (scene/bvh.cpp:54-222) finds the same nearest hits -- only exact ties in t could differ.  One thread per pixel, FP64, in the
// reference's operation order (--fmad=false), so the radiance agrees with the compiled reference to ~1e-15.
// The COLLADA loader, the GUI and the tile pool stay out of scope: the scene comes in as plain arrays (lfb_scene).
#include <algorithm>
#include <vector>

#include "lfb_internal.h"

namespace lfb {

namespace {

struct BvhNode {
  double lo[3], hi[3];
  int first, count;  // count > 0: leaf over prim_idx[first .. first + count); count == 0: interior, children first and first + 1
  int pad[2];
};

struct DevScene {
  const double* tri_pos;   // [nt][3][3]
  const double* tri_nrm;   // [nt][3][3]
  const int* tri_mat;
  const double* sph;       // [ns][4]
  const int* sph_mat;
  const double* mats;      // [nm][6]
  const double* lights;    // [nl][7]
  const BvhNode* nodes;
  const int* prim_idx;     // >= 0: triangle; < 0: sphere ~idx
  int nt, ns, nl, n_nodes;
};

struct Ray {
  double o[3], d[3], inv[3], min_t, max_t;
};
struct Hit {
  double t, n[3];
  int mat;
};

__device__ __forceinline__ double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
__device__ __forceinline__ void cross3(const double* a, const double* b, double* c) {
  c[0] = a[1] * b[2] - a[2] * b[1]; c[1] = a[2] * b[0] - a[0] * b[2]; c[2] = a[0] * b[1] - a[1] * b[0];
}
__device__ __forceinline__ void unit3(const double* a, double* u) {  // vector3D.h:215-218: times the reciprocal norm
  const double r = 1. / sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
  u[0] = a[0] * r; u[1] = a[1] * r; u[2] = a[2] * r;
}

// scene/triangle.cpp:26-113
__device__ __forceinline__ bool tri_hit(const double* __restrict__ P, const double* __restrict__ N, Ray& r, Hit* h) {
  double e1[3], e2[3], s[3], s1[3], s2[3];
#pragma unroll
  for (int a = 0; a < 3; a++) { e1[a] = P[3 + a] - P[a]; e2[a] = P[6 + a] - P[a]; s[a] = r.o[a] - P[a]; }
  cross3(r.d, e2, s1);
  cross3(s, e1, s2);
  const double den = dot3(s1, e1);
  const double t = dot3(s2, e2) / den, b1 = dot3(s1, s) / den, b2 = dot3(s2, r.d) / den;
  if (t < r.min_t || t > r.max_t) return false;
  if (b1 < 0 || b1 > 1) return false;
  if (b2 < 0 || b2 > 1) return false;
  if (b1 + b2 > 1) return false;
  if (!(t == t) || !(b1 == b1) || !(b2 == b2)) return false;
  if (!h) return true;
  const double b0 = 1 - b1 - b2;
  double n[3];
#pragma unroll
  for (int a = 0; a < 3; a++) n[a] = b0 * N[a] + b1 * N[3 + a] + b2 * N[6 + a];
  r.max_t = t;
  h->t = t;
  unit3(n, h->n);
  return true;
}

// scene/sphere.cpp:11-108
__device__ __forceinline__ bool sph_hit(const double* __restrict__ S, Ray& r, Hit* h) {
  const double oc[3] = {r.o[0] - S[0], r.o[1] - S[1], r.o[2] - S[2]};
  const double a = dot3(r.d, r.d), b = 2 * dot3(oc, r.d), c = dot3(oc, oc) - S[3] * S[3];
  double t1;
  if (b * b < 4.0 * a * c) return false;
  if (b * b == 4.0 * a * c) {
    const double root = (-b) / (2.0 * a);
    if (root < r.min_t || root > r.max_t) return false;
    t1 = root;
  } else {
    const double q = sqrt(b * b - 4.0 * a * c);
    const double r1 = (-b - q) / (2.0 * a), r2 = (-b + q) / (2.0 * a);
    const double lo = r1 < r2 ? r1 : r2, hi = r1 < r2 ? r2 : r1;
    if (lo > r.max_t || hi < r.min_t) return false;
    if (lo < r.min_t) {
      if (hi > r.max_t) return false;
      t1 = hi;
    } else {
      t1 = lo;
    }
  }
  if (!h) return true;
  r.max_t = t1;
  h->t = t1;
  const double p[3] = {r.o[0] + t1 * r.d[0] - S[0], r.o[1] + t1 * r.d[1] - S[1], r.o[2] + t1 * r.d[2] - S[2]};
  unit3(p, h->n);
  return true;
}

__device__ __forceinline__ bool box_hit(const BvhNode& n, const Ray& r) {
  double t0 = r.min_t, t1 = r.max_t;
#pragma unroll
  for (int a = 0; a < 3; a++) {
    double ta = (n.lo[a] - r.o[a]) * r.inv[a], tb = (n.hi[a] - r.o[a]) * r.inv[a];
    if (ta > tb) { const double q = ta; ta = tb; tb = q; }
    // NaN (0 * inf: the origin on a slab's plane of an axis-parallel ray) must not reject the box
    t0 = ta > t0 ? ta : t0;
    t1 = tb < t1 ? tb : t1;
  }
  return t0 <= t1;
}

// nearest hit (h != nullptr) or any hit (h == nullptr) in [r.min_t, r.max_t]
__device__ bool scene_hit(const DevScene& S, Ray& r, Hit* h) {
  if (S.n_nodes == 0) return false;
  int stack[64];
  int sp = 0;
  stack[sp++] = 0;
  bool hit = false;
  while (sp > 0) {
    const BvhNode& n = S.nodes[stack[--sp]];
    if (!box_hit(n, r)) continue;
    if (n.count > 0) {
      for (int q = 0; q < n.count; q++) {
        const int id = S.prim_idx[n.first + q];
        bool got;
        if (id >= 0) {
          got = tri_hit(S.tri_pos + 9 * (size_t)id, S.tri_nrm + 9 * (size_t)id, r, h);
          if (got && h) h->mat = S.tri_mat[id];
        } else {
          got = sph_hit(S.sph + 4 * (size_t)(~id), r, h);
          if (got && h) h->mat = S.sph_mat[~id];
        }
        if (got) {
          if (!h) return true;
          hit = true;
        }
      }
    } else if (sp + 2 <= 64) {
      stack[sp++] = n.first + 1;
      stack[sp++] = n.first;
    }
  }
  return hit;
}

struct SceneFrame {
  double pos[3], c2w[9], ex, ey, nclip, fclip;
  int W, H;
};

__global__ void __launch_bounds__(128) scene_kernel(DevScene S, SceneFrame F, char* __restrict__ out, size_t stride, int elem, int additive) {
  const int x = blockIdx.x * 16 + (threadIdx.x & 15), y = blockIdx.y * 8 + (threadIdx.x >> 4);
  if (x >= F.W || y >= F.H) return;
  double L[3] = {0.0, 0.0, 0.0};
  // Camera::generate_ray (camera.cpp:278-305)
  const double cx = F.ex * (2 * ((x + 0.5) / (double)F.W) - 1), cy = F.ey * (2 * ((y + 0.5) / (double)F.H) - 1);
  const double dcam[3] = {cx, cy, -1};
  double du[3];
  unit3(dcam, du);
  Ray r;
#pragma unroll
  for (int a = 0; a < 3; a++) {
    r.o[a] = F.pos[a];
    r.d[a] = F.c2w[3 * a] * du[0] + F.c2w[3 * a + 1] * du[1] + F.c2w[3 * a + 2] * du[2];
    r.inv[a] = 1.0 / r.d[a];
  }
  r.min_t = F.nclip; r.max_t = F.fclip;
  Hit h;
  if (scene_hit(S, r, &h)) {
    const double* M = S.mats + 6 * h.mat;
    const bool emissive = M[3] > 0 || M[4] > 0 || M[5] > 0;
    if (emissive) { L[0] = M[3]; L[1] = M[4]; L[2] = M[5]; }  // zero bounce
    // make_coord_space (pathtracer/bsdf.cpp:20-43)
    double z[3] = {h.n[0], h.n[1], h.n[2]}, hh[3] = {h.n[0], h.n[1], h.n[2]}, xx[3], yy[3];
    if (fabs(hh[0]) <= fabs(hh[1]) && fabs(hh[0]) <= fabs(hh[2])) hh[0] = 1.0;
    else if (fabs(hh[1]) <= fabs(hh[0]) && fabs(hh[1]) <= fabs(hh[2])) hh[1] = 1.0;
    else hh[2] = 1.0;
    { const double nz = sqrt(z[0] * z[0] + z[1] * z[1] + z[2] * z[2]); z[0] /= nz; z[1] /= nz; z[2] /= nz; }
    cross3(hh, z, yy);
    { const double ny = sqrt(yy[0] * yy[0] + yy[1] * yy[1] + yy[2] * yy[2]); yy[0] /= ny; yy[1] /= ny; yy[2] /= ny; }
    cross3(z, yy, xx);
    { const double nx = sqrt(xx[0] * xx[0] + xx[1] * xx[1] + xx[2] * xx[2]); xx[0] /= nx; xx[1] /= nx; xx[2] /= nx; }
    const double hit_p[3] = {r.o[0] + r.d[0] * h.t, r.o[1] + r.d[1] * h.t, r.o[2] + r.d[2] * h.t};
    double Ld[3] = {0.0, 0.0, 0.0};
    const double pi = 3.14159265358979323846264338327950288;
    const float eps_f = 0.00001f;  // EPS_F, CGL/include/CGL/misc.h:13
    for (int l = 0; l < S.nl; l++) {
      const double* A = S.lights + 7 * l;
      double wi[3], dist;
      if (A[0] == 0) {  // DirectionalLight (scene/light.cpp:11-24): dirToLight = -lightDir.unit()
        double u[3];
        unit3(A + 4, u);
        wi[0] = -u[0]; wi[1] = -u[1]; wi[2] = -u[2];
        dist = INFINITY;
      } else {  // PointLight (:49-60)
        const double d[3] = {A[4] - hit_p[0], A[5] - hit_p[1], A[6] - hit_p[2]};
        unit3(d, wi);
        dist = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
      }
      const double wo[3] = {dot3(xx, wi), dot3(yy, wi), dot3(z, wi)};  // w2o * wi
      if (wo[2] < 0) continue;  // the light is behind the surface (:186-188)
      Ray sh;
#pragma unroll
      for (int a = 0; a < 3; a++) { sh.o[a] = hit_p[a]; sh.d[a] = wi[a]; sh.inv[a] = 1.0 / wi[a]; }
      sh.min_t = eps_f; sh.max_t = dist - eps_f;
      if (scene_hit(S, sh, nullptr)) continue;  // occluded
      double wu[3];
      unit3(wo, wu);
      if (!emissive) {
#pragma unroll
        for (int c = 0; c < 3; c++) Ld[c] += (((1.0 / pi) * M[c]) * A[1 + c] * wu[2]) / 1.0;
      }
    }
    if (S.nl > 0) {
#pragma unroll
      for (int c = 0; c < 3; c++) L[c] = L[c] + Ld[c] / (double)S.nl;
    }
  }
  const size_t p = (size_t)x + (size_t)y * F.W;
  if (elem == LFB_F32x3) {
    float* o = reinterpret_cast<float*>(out + p * stride);
    if (additive) { o[0] += (float)L[0]; o[1] += (float)L[1]; o[2] += (float)L[2]; }
    else { o[0] = (float)L[0]; o[1] = (float)L[1]; o[2] = (float)L[2]; }
  } else {
    double* o = reinterpret_cast<double*>(out + p * stride);
    if (additive) { o[0] += L[0]; o[1] += L[1]; o[2] += L[2]; }
    else { o[0] = L[0]; o[1] = L[1]; o[2] = L[2]; }
  }
}

// ---- host: BVH build ------------------------------------------------------------------------------------------------
struct PrimBox {
  double lo[3], hi[3], c[3];
  int id;
};

void build_node(std::vector<BvhNode>& nodes, std::vector<PrimBox>& prims, int node, int first, int count) {
  BvhNode n;
  for (int a = 0; a < 3; a++) { n.lo[a] = INFINITY; n.hi[a] = -INFINITY; }
  double clo[3] = {INFINITY, INFINITY, INFINITY}, chi[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (int q = first; q < first + count; q++)
    for (int a = 0; a < 3; a++) {
      n.lo[a] = std::min(n.lo[a], prims[q].lo[a]); n.hi[a] = std::max(n.hi[a], prims[q].hi[a]);
      clo[a] = std::min(clo[a], prims[q].c[a]); chi[a] = std::max(chi[a], prims[q].c[a]);
    }
  for (int a = 0; a < 3; a++) {  // a hair of slack: the slab test must never reject a hit the primitive test accepts
    const double pad = 1e-9 * std::max(1.0, std::max(fabs(n.lo[a]), fabs(n.hi[a])));
    n.lo[a] -= pad; n.hi[a] += pad;
  }
  n.pad[0] = n.pad[1] = 0;
  if (count <= 4) {
    n.first = first; n.count = count;
    nodes[node] = n;
    return;
  }
  int axis = 0;
  for (int a = 1; a < 3; a++)
    if (chi[a] - clo[a] > chi[axis] - clo[axis]) axis = a;
  const int mid = first + count / 2;
  std::nth_element(prims.begin() + first, prims.begin() + mid, prims.begin() + first + count,
                   [axis](const PrimBox& p, const PrimBox& q) { return p.c[axis] < q.c[axis] || (p.c[axis] == q.c[axis] && p.id < q.id); });
  const int left = (int)nodes.size();
  nodes.push_back(BvhNode());
  nodes.push_back(BvhNode());
  n.first = left; n.count = 0;
  nodes[node] = n;
  build_node(nodes, prims, left, first, mid - first);
  build_node(nodes, prims, left + 1, mid, first + count - mid);
}

}  // namespace

struct SceneStore {
  double *tri_pos = nullptr, *tri_nrm = nullptr, *sph = nullptr, *mats = nullptr, *lights = nullptr;
  int *tri_mat = nullptr, *sph_mat = nullptr, *prim_idx = nullptr;
  BvhNode* nodes = nullptr;
  int nt = 0, ns = 0, nm = 0, nl = 0, n_nodes = 0;
};

void scene_free(SceneStore* s) {
  if (!s) return;
  cudaFree(s->tri_pos); cudaFree(s->tri_nrm); cudaFree(s->sph); cudaFree(s->mats); cudaFree(s->lights);
  cudaFree(s->tri_mat); cudaFree(s->sph_mat); cudaFree(s->prim_idx); cudaFree(s->nodes);
  delete s;
}

template <typename T>
static cudaError_t upload(T** dst, const T* src, size_t n) {
  *dst = nullptr;
  if (n == 0) return cudaSuccess;
  cudaError_t err = cudaMalloc((void**)dst, sizeof(T) * n);
  if (err != cudaSuccess) return err;
  return cudaMemcpy(*dst, src, sizeof(T) * n, cudaMemcpyHostToDevice);
}

// Validates the arrays, builds the BVH on the host and uploads everything.  err_msg: a static string on LFB_ERR_INVALID.
cudaError_t scene_upload(const lfb_scene* sc, SceneStore** out, const char** err_msg) {
  *out = nullptr;
  *err_msg = nullptr;
  for (int t = 0; t < sc->n_tri; t++)
    if (sc->tri_mat[t] < 0 || sc->tri_mat[t] >= sc->n_mat) { *err_msg = "a triangle's material index is out of range"; return cudaSuccess; }
  for (int k = 0; k < sc->n_sph; k++)
    if (sc->sph_mat[k] < 0 || sc->sph_mat[k] >= sc->n_mat) { *err_msg = "a sphere's material index is out of range"; return cudaSuccess; }
  for (int l = 0; l < sc->n_lights; l++)
    if (sc->lights[7 * l] != 0.0 && sc->lights[7 * l] != 1.0) { *err_msg = "light kind must be 0 (directional) or 1 (point): area lights are not supported"; return cudaSuccess; }
  std::vector<PrimBox> prims;
  prims.reserve((size_t)sc->n_tri + sc->n_sph);
  for (int t = 0; t < sc->n_tri; t++) {
    PrimBox b;
    const double* P = sc->tri_pos + 9 * (size_t)t;
    for (int a = 0; a < 3; a++) {
      b.lo[a] = std::min(P[a], std::min(P[3 + a], P[6 + a]));
      b.hi[a] = std::max(P[a], std::max(P[3 + a], P[6 + a]));
      b.c[a] = 0.5 * (b.lo[a] + b.hi[a]);
    }
    b.id = t;
    prims.push_back(b);
  }
  for (int k = 0; k < sc->n_sph; k++) {
    PrimBox b;
    const double* S = sc->spheres + 4 * (size_t)k;
    for (int a = 0; a < 3; a++) { b.lo[a] = S[a] - fabs(S[3]); b.hi[a] = S[a] + fabs(S[3]); b.c[a] = S[a]; }
    b.id = ~k;
    prims.push_back(b);
  }
  std::vector<BvhNode> nodes;
  if (!prims.empty()) {
    nodes.reserve(2 * prims.size());
    nodes.push_back(BvhNode());
    build_node(nodes, prims, 0, 0, (int)prims.size());
  }
  std::vector<int> idx(prims.size());
  for (size_t q = 0; q < prims.size(); q++) idx[q] = prims[q].id;
  SceneStore* s = new SceneStore();
  s->nt = sc->n_tri; s->ns = sc->n_sph; s->nm = sc->n_mat; s->nl = sc->n_lights; s->n_nodes = (int)nodes.size();
  cudaError_t err = upload(&s->tri_pos, sc->tri_pos, 9 * (size_t)sc->n_tri);
  if (err == cudaSuccess) err = upload(&s->tri_nrm, sc->tri_nrm, 9 * (size_t)sc->n_tri);
  if (err == cudaSuccess) err = upload(&s->tri_mat, sc->tri_mat, (size_t)sc->n_tri);
  if (err == cudaSuccess) err = upload(&s->sph, sc->spheres, 4 * (size_t)sc->n_sph);
  if (err == cudaSuccess) err = upload(&s->sph_mat, sc->sph_mat, (size_t)sc->n_sph);
  if (err == cudaSuccess) err = upload(&s->mats, sc->materials, 6 * (size_t)sc->n_mat);
  if (err == cudaSuccess) err = upload(&s->lights, sc->lights, 7 * (size_t)sc->n_lights);
  if (err == cudaSuccess) err = upload(&s->nodes, nodes.data(), nodes.size());
  if (err == cudaSuccess) err = upload(&s->prim_idx, idx.data(), idx.size());
  if (err != cudaSuccess) { scene_free(s); return err; }
  *out = s;
  return cudaSuccess;
}

cudaError_t launch_scene(const SceneStore* s, const lfb_camera* cam, int W, int H, void* out, size_t stride, int elem, int additive, cudaStream_t st) {
  DevScene D;
  D.tri_pos = s->tri_pos; D.tri_nrm = s->tri_nrm; D.tri_mat = s->tri_mat; D.sph = s->sph; D.sph_mat = s->sph_mat;
  D.mats = s->mats; D.lights = s->lights; D.nodes = s->nodes; D.prim_idx = s->prim_idx;
  D.nt = s->nt; D.ns = s->ns; D.nl = s->nl; D.n_nodes = s->n_nodes;
  SceneFrame F;
  for (int a = 0; a < 3; a++) F.pos[a] = cam->pos[a];
  for (int a = 0; a < 9; a++) F.c2w[a] = cam->c2w[a];
  const double pi = 3.14159265358979323846264338327950288;  // CGL's PI
  F.ex = tan(0.5 * (cam->hfov_deg * (pi / 180.0)));
  F.ey = tan(0.5 * (cam->vfov_deg * (pi / 180.0)));
  F.nclip = cam->nclip; F.fclip = cam->fclip;
  F.W = W; F.H = H;
  scene_kernel<<<dim3((W + 15) / 16, (H + 7) / 8), 128, 0, st>>>(D, F, (char*)out, stride, elem, additive);
  return cudaGetLastError();
}

}  // namespace lfb
