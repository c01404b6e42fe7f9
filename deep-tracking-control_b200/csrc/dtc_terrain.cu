// terrain rasterisation on the device.  The reference builds its int16 heightfield on the host with
// isaacgym.terrain_utils (legged_gym/utils/terrain.py:9-243: pyramid stairs, discrete obstacles, stepping stones laid out as
// num_rows x num_cols sub-terrains inside a flat border) and uploads it; here the host only draws the handful of random
// parameters (as the reference does, with numpy) and one thread per map cell evaluates the sub-terrain's closed form, so
// the 1760 x 1120 map, the env-origin heights and the foothold kernel's tables never exist on the host.
// Oracle: deep-tracking-control_b200/sim_stub.make_heightmap (numpy), bit-exact (tests/test_env_gpu.py).
#include "dtc_common.cuh"

#define TERRAIN_TABLE 256  // int32 entries per sub-terrain: stepping-stone column offsets or 20 x {x, y, w, l, height} rectangles

__device__ __forceinline__ int terrain_cell(const dtc_subterrain& s, const int32_t* __restrict__ tab, int x, int y, int px) {
  int h = 0;
  if (s.type == 1) {  // stepping stones: stones of side a separated by gaps b over a pit of depth c; column k starts at tab[k]
    const int P = s.a + s.b, cx = x / P;
    h = s.c;
    if (x - cx * P < s.a) {
      const int y0 = tab[cx];
      if (y < max(0, y0 - s.b) || (y >= y0 && (y - y0) % P < s.a)) h = 0;
    }
  } else if (s.type == 2) {  // pyramid stairs: step width a, step height b (signed), platform c
    const int m = min(min(x, px - 1 - x), min(y, px - 1 - y)) / s.a;
    int T = 0;  // number of rings the host loop lays down: ring t while px - 2 a (t - 1) > c
    if (px > s.c) T = (px - s.c - 1) / (2 * s.a) + 1;
    h = s.b * min(m, T);
  } else if (s.type == 3) {  // discrete obstacles: a rectangles, later ones override earlier ones
    for (int r = 0; r < s.a; ++r) {
      const int32_t* q = tab + 5 * r;
      if (x >= q[0] && x < q[0] + q[2] && y >= q[1] && y < q[1] + q[3]) h = q[4];
    }
  }
  if (s.platform_half > 0) {
    const int c = px / 2, p = s.platform_half;
    if (x >= c - p && x < c + p && y >= c - p && y < c + p) h = 0;
  }
  return h;
}

__global__ void __launch_bounds__(256) k_terrain_rasterize(int rows, int cols, int border_px, int sub_px, int n_rows, int n_cols,
                                                           const dtc_subterrain* __restrict__ subs, const int32_t* __restrict__ tables,
                                                           int16_t* __restrict__ out) {
  const int64_t total = (int64_t)rows * cols;
  for (int64_t e = blockIdx.x * 256ll + threadIdx.x; e < total; e += gridDim.x * 256ll) {
    const int gx = (int)(e / cols), gy = (int)(e - (int64_t)gx * cols);
    const int lx = gx - border_px, ly = gy - border_px;
    int h = 0;
    if (lx >= 0 && ly >= 0 && lx < n_rows * sub_px && ly < n_cols * sub_px) {
      const int i = lx / sub_px, j = ly / sub_px, s = i * n_cols + j;
      h = terrain_cell(subs[s], tables + (size_t)s * TERRAIN_TABLE, lx - i * sub_px, ly - j * sub_px, sub_px);
    }
    out[e] = (int16_t)h;
  }
}

// origins[i][j] = ((i + 0.5) L, (j + 0.5) L, max over the central 20 x 20 cells * vertical_scale): one warp per sub-terrain
__global__ void __launch_bounds__(32) k_terrain_origins(int cols, int border_px, int sub_px, int n_cols, double terrain_length,
                                                        double vertical_scale, const int16_t* __restrict__ map, float* __restrict__ origins) {
  const int s = blockIdx.x, i = s / n_cols, j = s - i * n_cols, c = sub_px / 2;
  int mx = -32768;
  for (int e = threadIdx.x; e < 400; e += 32) {
    const int x = border_px + i * sub_px + c - 10 + e / 20, y = border_px + j * sub_px + c - 10 + e % 20;
    mx = max(mx, (int)map[(size_t)x * cols + y]);
  }
  for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (threadIdx.x == 0) {
    origins[s * 3 + 0] = (float)(((double)i + 0.5) * terrain_length);
    origins[s * 3 + 1] = (float)(((double)j + 0.5) * terrain_length);
    origins[s * 3 + 2] = (float)((double)mx * vertical_scale);
  }
}

extern "C" int dtc_terrain_rasterize(int32_t rows, int32_t cols, int32_t border_px, int32_t sub_px, int32_t n_rows, int32_t n_cols,
                                     const dtc_subterrain* subs, const int32_t* tables, double terrain_length, double vertical_scale,
                                     int16_t* height_samples, float* terrain_origins, void* stream) {
  DTC_NVTX("dtc_terrain_rasterize");
  if (!subs || !tables || !height_samples) DTC_FAIL(DTC_ERR_ARG, "dtc_terrain_rasterize: null argument");
  if (rows <= 0 || cols <= 0 || sub_px < 20 || n_rows <= 0 || n_cols <= 0 || border_px < 0 || n_rows * sub_px + 2 * border_px > rows ||
      n_cols * sub_px + 2 * border_px > cols)
    DTC_FAIL(DTC_ERR_ARG, "dtc_terrain_rasterize: %d x %d sub-terrains of %d px + 2 x %d px border do not fit a %d x %d map", n_rows, n_cols,
             sub_px, border_px, rows, cols);
  cudaStream_t st = (cudaStream_t)stream;
  k_terrain_rasterize<<<148 * 8, 256, 0, st>>>(rows, cols, border_px, sub_px, n_rows, n_cols, subs, tables, height_samples);
  DTC_CHECK_LAUNCH("k_terrain_rasterize");
  if (terrain_origins) {
    k_terrain_origins<<<n_rows * n_cols, 32, 0, st>>>(cols, border_px, sub_px, n_cols, terrain_length, vertical_scale, height_samples,
                                                      terrain_origins);
    DTC_CHECK_LAUNCH("k_terrain_origins");
  }
  return DTC_OK;
}

// ------------------------------------------------------------------ generic painter: ordered rectangles
// Every generator of legged_gym/utils/terrain.py (and of isaacgym.terrain_utils behind it) that the DTC tasks use is a sequence of
// `height_field_raw[x0:x1, y0:y1] = h` assignments on a constant background: nested squares (pyramid stairs), random boxes
// (discrete obstacles), stones (stepping stones, stones_everywhere), one or two boxes (pit, gap).  The host replays the generator's
// loops and numpy draws and records the assignments instead of executing them; one thread per map cell then takes the LAST
// rectangle that covers it (later assignments overwrite earlier ones), or the sub-terrain's background.
__global__ void __launch_bounds__(256) k_terrain_paint(int rows, int cols, int border_px, int len_px, int wid_px, int n_rows, int n_cols,
                                                       const dtc_subterrain* __restrict__ subs, const int32_t* __restrict__ rects,
                                                       int16_t* __restrict__ out) {
  const int64_t total = (int64_t)rows * cols;
  for (int64_t e = blockIdx.x * 256ll + threadIdx.x; e < total; e += gridDim.x * 256ll) {
    const int gx = (int)(e / cols), gy = (int)(e - (int64_t)gx * cols);
    const int lx = gx - border_px, ly = gy - border_px;
    int h = 0;
    if (lx >= 0 && ly >= 0 && lx < n_rows * len_px && ly < n_cols * wid_px) {
      const int i = lx / len_px, j = ly / wid_px;
      const dtc_subterrain s = subs[i * n_cols + j];  // type 5: a = number of rectangles, b = index of the first, c = background
      const int x = lx - i * len_px, y = ly - j * wid_px;
      h = s.c;
      for (int r = s.a - 1; r >= 0; --r) {
        const int32_t* q = rects + 5 * (size_t)(s.b + r);
        if (x >= q[0] && x < q[1] && y >= q[2] && y < q[3]) { h = q[4]; break; }
      }
    }
    out[e] = (int16_t)h;
  }
}

// env origins of add_terrain_to_map (terrain.py:152-160): centre of the sub-terrain, height = max over the window [x1,x2) x [y1,y2)
__global__ void __launch_bounds__(128) k_terrain_origins_win(int cols, int border_px, int len_px, int wid_px, int n_cols, double terrain_length,
                                                             double terrain_width, double vertical_scale, int x1, int x2, int y1, int y2,
                                                             const int16_t* __restrict__ map, float* __restrict__ origins) {
  __shared__ int red[4];
  const int s = blockIdx.x, i = s / n_cols, j = s - i * n_cols;
  const int w = y2 - y1, n = (x2 - x1) * w;
  int mx = -32768;
  for (int e = threadIdx.x; e < n; e += 128) {
    const int x = border_px + i * len_px + x1 + e / w, y = border_px + j * wid_px + y1 + e % w;
    mx = max(mx, (int)map[(size_t)x * cols + y]);
  }
  for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    mx = max(max(red[0], red[1]), max(red[2], red[3]));
    origins[s * 3 + 0] = (float)(((double)i + 0.5) * terrain_length);
    origins[s * 3 + 1] = (float)(((double)j + 0.5) * terrain_width);
    origins[s * 3 + 2] = (float)((double)mx * vertical_scale);
  }
}

extern "C" int dtc_terrain_paint(int32_t rows, int32_t cols, int32_t border_px, int32_t len_px, int32_t wid_px, int32_t n_rows, int32_t n_cols,
                                 const dtc_subterrain* subs, const int32_t* rects, const int32_t origin_window[4], double terrain_length,
                                 double terrain_width, double vertical_scale, int16_t* height_samples, float* terrain_origins, void* stream) {
  DTC_NVTX("dtc_terrain_paint");
  if (!subs || !height_samples) DTC_FAIL(DTC_ERR_ARG, "dtc_terrain_paint: null argument");
  if (rows <= 0 || cols <= 0 || len_px <= 0 || wid_px <= 0 || n_rows <= 0 || n_cols <= 0 || border_px < 0 || n_rows * len_px + 2 * border_px > rows ||
      n_cols * wid_px + 2 * border_px > cols)
    DTC_FAIL(DTC_ERR_ARG, "dtc_terrain_paint: %d x %d sub-terrains of %d x %d px + 2 x %d px border do not fit a %d x %d map", n_rows, n_cols,
             len_px, wid_px, border_px, rows, cols);
  cudaStream_t st = (cudaStream_t)stream;
  // this library carries its own (static) CUDA runtime; when a launch is the first thing it is asked to do, the runtime initialises
  // inside the launch path and compute-sanitizer reports the benign first-try cuKernelGetFunction error -> initialise it up front
  static const cudaError_t runtime_ready = cudaFree(nullptr);
  (void)runtime_ready;
  k_terrain_paint<<<148 * 8, 256, 0, st>>>(rows, cols, border_px, len_px, wid_px, n_rows, n_cols, subs, rects, height_samples);
  DTC_CHECK_LAUNCH("k_terrain_paint");
  if (terrain_origins) {
    if (!origin_window || origin_window[0] < 0 || origin_window[1] > len_px || origin_window[2] < 0 || origin_window[3] > wid_px ||
        origin_window[0] >= origin_window[1] || origin_window[2] >= origin_window[3])
      DTC_FAIL(DTC_ERR_ARG, "dtc_terrain_paint: bad origin window");
    k_terrain_origins_win<<<n_rows * n_cols, 128, 0, st>>>(cols, border_px, len_px, wid_px, n_cols, terrain_length, terrain_width, vertical_scale,
                                                           origin_window[0], origin_window[1], origin_window[2], origin_window[3],
                                                           height_samples, terrain_origins);
    DTC_CHECK_LAUNCH("k_terrain_origins_win");
  }
  return DTC_OK;
}
