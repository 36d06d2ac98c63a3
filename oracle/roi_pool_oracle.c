/* ORACLE (test infrastructure, NOT product code): plain C restatement of the
 * reference's CPU ROI max pooling, nms_net/roi_pooling_layer/roi_pooling_op.cc
 *   forward  :128-187  (per output element (n, ph, pw, c): C round() of the
 *            scaled roi corners, malformed rois forced to 1x1, floor/ceil bin
 *            edges in float, clip to the map, empty bin -> 0 / argmax -1,
 *            strict '>' so the first maximum wins, argmax index
 *            (h*W + w)*C + c inside the roi's image)
 *   backward :374-449  (per input element: rois in index order, then ph, pw
 *            ascending; float accumulation in that order)
 * Build: make -C oracle.  Checked against the reference's own object code
 * (oracle/_ref/libroi_pool_ref.so) by tests/test_roi_pool_oracle.py. */
#include <float.h>
#include <math.h>
#include <stdint.h>

static int imin(int a, int b) { return a < b ? a : b; }
static int imax(int a, int b) { return a > b ? a : b; }

int oracle_roi_pool_fwd(const float* data, int batch, int height, int width, int channels,
                        const float* rois, int num_rois, int pooled_h, int pooled_w,
                        float spatial_scale, float* top, int32_t* argmax) {
  (void)batch;
  const int64_t total = (int64_t)num_rois * pooled_h * pooled_w * channels;
  for (int64_t b = 0; b < total; ++b) {
    int64_t n = b;
    const int c = (int)(n % channels); n /= channels;
    const int pw = (int)(n % pooled_w); n /= pooled_w;
    const int ph = (int)(n % pooled_h); n /= pooled_h;
    const float* roi = rois + n * 5;
    const int roi_batch_ind = (int)roi[0];
    const int roi_start_w = (int)round(roi[1] * spatial_scale);
    const int roi_start_h = (int)round(roi[2] * spatial_scale);
    const int roi_end_w = (int)round(roi[3] * spatial_scale);
    const int roi_end_h = (int)round(roi[4] * spatial_scale);
    const int roi_width = imax(roi_end_w - roi_start_w + 1, 1);
    const int roi_height = imax(roi_end_h - roi_start_h + 1, 1);
    const float bin_size_h = (float)roi_height / (float)pooled_h;
    const float bin_size_w = (float)roi_width / (float)pooled_w;
    int hstart = (int)floorf(ph * bin_size_h);
    int wstart = (int)floorf(pw * bin_size_w);
    int hend = (int)ceilf((ph + 1) * bin_size_h);
    int wend = (int)ceilf((pw + 1) * bin_size_w);
    hstart = imin(imax(hstart + roi_start_h, 0), height);
    hend = imin(imax(hend + roi_start_h, 0), height);
    wstart = imin(imax(wstart + roi_start_w, 0), width);
    wend = imin(imax(wend + roi_start_w, 0), width);
    const int is_empty = (hend <= hstart) || (wend <= wstart);
    float maxval = is_empty ? 0.f : -FLT_MAX;
    int maxidx = -1;
    const float* img = data + (int64_t)roi_batch_ind * channels * height * width;
    for (int h = hstart; h < hend; ++h)
      for (int w = wstart; w < wend; ++w) {
        const int idx = (h * width + w) * channels + c;
        if (img[idx] > maxval) { maxval = img[idx]; maxidx = idx; }
      }
    top[b] = maxval;
    argmax[b] = maxidx;
  }
  return 0;
}

int oracle_roi_pool_bwd(int batch, int height, int width, int channels, const float* rois,
                        int num_rois, const int32_t* argmax, const float* top_diff, int pooled_h,
                        int pooled_w, float spatial_scale, float* bottom_diff) {
  const int64_t total = (int64_t)batch * height * width * channels;
  for (int64_t b = 0; b < total; ++b) {
    int64_t n = b;
    const int c = (int)(n % channels); n /= channels;
    const int w = (int)(n % width); n /= width;
    const int h = (int)(n % height); n /= height;
    float gradient = 0.f;
    for (int roi_n = 0; roi_n < num_rois; ++roi_n) {
      const float* roi = rois + (int64_t)roi_n * 5;
      if (n != (int)roi[0]) continue;
      const int roi_start_w = (int)round(roi[1] * spatial_scale);
      const int roi_start_h = (int)round(roi[2] * spatial_scale);
      const int roi_end_w = (int)round(roi[3] * spatial_scale);
      const int roi_end_h = (int)round(roi[4] * spatial_scale);
      if (!(w >= roi_start_w && w <= roi_end_w && h >= roi_start_h && h <= roi_end_h)) continue;
      const int64_t offset = (int64_t)roi_n * pooled_h * pooled_w * channels;
      const int roi_width = imax(roi_end_w - roi_start_w + 1, 1);
      const int roi_height = imax(roi_end_h - roi_start_h + 1, 1);
      const float bin_size_h = (float)roi_height / (float)pooled_h;
      const float bin_size_w = (float)roi_width / (float)pooled_w;
      int phstart = (int)floorf((float)(h - roi_start_h) / bin_size_h);
      int phend = (int)ceilf((float)(h - roi_start_h + 1) / bin_size_h);
      int pwstart = (int)floorf((float)(w - roi_start_w) / bin_size_w);
      int pwend = (int)ceilf((float)(w - roi_start_w + 1) / bin_size_w);
      phstart = imin(imax(phstart, 0), pooled_h);
      phend = imin(imax(phend, 0), pooled_h);
      pwstart = imin(imax(pwstart, 0), pooled_w);
      pwend = imin(imax(pwend, 0), pooled_w);
      for (int ph = phstart; ph < phend; ++ph)
        for (int pw = pwstart; pw < pwend; ++pw)
          if (argmax[offset + (ph * pooled_w + pw) * channels + c] == (h * width + w) * channels + c)
            gradient += top_diff[offset + (ph * pooled_w + pw) * channels + c];
    }
    bottom_diff[b] = gradient;
  }
  return 0;
}
