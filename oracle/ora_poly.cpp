// ora_poly.cpp - CPU oracle: the poly.cpp pipeline (poly.cpp:51-131, config 1) replayed with the same twelve
// buffers (mem0..mem9 planes, memBig, memLS) and the same aliasing.  TEST INFRASTRUCTURE ONLY (see rd_oracle.h).
#include <vector>
#include "ora_internal.h"

extern "C" void ora_poly_frame(const uint8_t *img, int ws, int iw, int ih, float minerror, int sizeThre, int strengthThre,
                               int32_t *lsIdOut, ora_ls_t *lsListOut, float *thinOut) {
  const size_t n = (size_t)iw * ih;
  std::vector<std::vector<int32_t>> mem(10, std::vector<int32_t>(n, 0));   // poly.cpp:75-85 : calloc / memset 0
  std::vector<int32_t> memBig(n * 4, 0), memLS(n * 4, 0);
  int32_t *m[10];
  for (int i = 0; i < 10; i++) m[i] = mem[i].data();
  memcpy(m[0], img, (size_t)ws * ih);                                       // poly.cpp:90

  // poly.cpp:104-116 : Stage A
  ora_convert_plab_bgr((uint32_t *)m[4], (const uint8_t *)m[0], iw, ih, ws);
  ora_unpack_f_f_f_plab((float *)m[1], (float *)m[2], (float *)m[3], (const uint32_t *)m[4], iw, ih);
  ora_iirblur_f_f((float *)m[0], (const float *)m[1], (float *)m[4], (float *)m[5], 2, iw, ih);
  ora_iirblur_f_f((float *)m[1], (const float *)m[2], (float *)m[4], (float *)m[5], 2, iw, ih);
  ora_iirblur_f_f((float *)m[2], (const float *)m[3], (float *)m[4], (float *)m[5], 2, iw, ih);
  ora_pack_plab_f_f_f((uint32_t *)m[4], (const float *)m[0], (const float *)m[1], (const float *)m[2], iw, ih);
  ora_edgevec_f2_f((float *)memBig.data(), (const float *)m[0], iw, ih);
  ora_edge_f_plab((float *)m[5], (const uint32_t *)m[4], iw, ih);
  ora_thinthres_f_f_f2((float *)m[2], (const float *)m[5], (const float *)memBig.data(), iw, ih);
  if (thinOut) memcpy(thinOut, m[2], n * 4);

  // poly.cpp:117-121 : edge cleanup
  ora_threshold_f_f((float *)m[9], (const float *)m[2], 0.0f, 0.0f, 1.0f, (int)n);
  ora_cast_i_f(m[8], (const float *)m[9], 1.0f, (int)n);
  ora::label8x(m[3], m[8], m[9], 0, iw, ih);
  ora_clear(m[4], iw * ih * 4);
  ora_calcStrength(m[4], (const float *)m[2], m[3], iw, ih);
  ora_filterStrength(m[3], m[4], strengthThre, iw, ih);
  ora_threshold_i_i(m[3], m[3], 0, 0, 1, (int)n);

  // poly.cpp:123
  ora_polyline_execute((ora_ls_t *)memLS.data(), iw * ih * 4 * 4, m[0], m[3], memBig.data(), m[4], m[5], m[6], m[7], m[8], m[9],
                       minerror, sizeThre, iw, ih, 0);
  memcpy(lsIdOut, m[0], n * 4);
  memcpy(lsListOut, memLS.data(), n * 16);
}
