// Closed-form index arithmetic of the 3-D shifted-window attention (host + device).
//
// Replaces the copy chain torch.roll(-ss) -> pad_3d -> window_partition_3d and its inverse
// (aurora/model/swin3d.py:470-503), maybe_adjust_windows (aurora/model/util.py:53-71) and the mask
// builder compute_3d_shifted_window_mask (aurora/model/swin3d.py:303-360) by pure addressing:
// (window, in-window token) -> source token of the un-rolled, un-padded (C,H,W) grid, or "pad".
#pragma once

#include <stdint.h>

namespace ab {

constexpr int kPadGroup = 27;  // group id of zero-padded tokens (swin3d.py:348-352)

struct WinGeom {
  int res[3];   // C, H, W of the token grid
  int ws[3];    // window, clamped to the resolution
  int ss[3];    // cyclic shift (0 on clamped axes)
  int lo[3];    // zero padding in FRONT of each axis (pad // 2); the rest goes to the back
  int nwin[3];  // windows per axis of the padded grid
  int ntok;     // tokens per window
  int nwindows; // windows per batch element
  int shifted;  // any(ss != 0): the group mask applies
  int warped;   // longitude wraps (left/right groups merged)
};

// Host-side construction from the configured window / shift.
inline WinGeom make_win_geom(const int res[3], const int ws0[3], const int ss0[3], int warped) {
  WinGeom g;
  g.ntok = 1;
  g.nwindows = 1;
  g.shifted = 0;
  for (int a = 0; a < 3; ++a) {
    g.res[a] = res[a];
    if (res[a] <= ws0[a]) {  // maybe_adjust_windows
      g.ws[a] = res[a];
      g.ss[a] = 0;
    } else {
      g.ws[a] = ws0[a];
      g.ss[a] = ss0[a];
    }
    int pad = (g.ws[a] - res[a] % g.ws[a]) % g.ws[a];
    g.lo[a] = pad / 2;
    g.nwin[a] = (res[a] + pad) / g.ws[a];
    g.ntok *= g.ws[a];
    g.nwindows *= g.nwin[a];
    if (g.ss[a] != 0) g.shifted = 1;
  }
  g.warped = warped;
  return g;
}

// Source token (flat (c*H + h)*W + w) of in-window token `tok` of window `win`, or -1 if the position
// is zero padding.  `group` receives the attention group id (only meaningful when g.shifted).
__host__ __device__ inline int win_source_token(const WinGeom& g, int win, int tok, int* group) {
  int k[3], i[3];
  k[2] = win % g.nwin[2];
  int t = win / g.nwin[2];
  k[1] = t % g.nwin[1];
  k[0] = t / g.nwin[1];
  i[2] = tok % g.ws[2];
  t = tok / g.ws[2];
  i[1] = t % g.ws[1];
  i[0] = t / g.ws[1];
  int src[3];
  int grp = 0;
  bool valid = true;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const int q = k[a] * g.ws[a] + i[a] - g.lo[a];  // coordinate in the shifted, un-padded frame
    valid = valid && (q >= 0) && (q < g.res[a]);
    int s = q + g.ss[a];
    if (s >= g.res[a]) s -= g.res[a];
    src[a] = s;
    int b;
    if (g.ss[a] == 0) {
      b = 2;
    } else {
      b = (q < g.res[a] - g.ws[a]) ? 0 : ((q < g.res[a] - g.ss[a]) ? 1 : 2);
    }
    if (a == 2 && g.warped && b == 1) b = 2;
    grp = grp * 3 + b;
  }
  if (!valid) {
    *group = kPadGroup;
    return -1;
  }
  *group = grp;
  return (src[0] * g.res[1] + src[1]) * g.res[2] + src[2];
}

}  // namespace ab
