// opencv2/opencv.hpp - a stand-in for the few OpenCV 2/3 names the reference's demo programs use (rect.cpp, poly.cpp,
// vidrect.cpp, vidpoly.cpp), so that those programs compile UNCHANGED in an image without OpenCV and run against
// librectdetect_b200.so (or against the reference itself, oracle/_ref/librd_ref.so).  TEST INFRASTRUCTURE only: it is not an
// image library.  What it does instead of OpenCV:
//   imread / imwrite            binary PPM (P6), converted to / from BGR rows with a 4-byte-aligned step
//   VideoCapture(path)          a raw stream file: "RDV1 <iw> <ih> <nframes>\n" followed by nframes x ih x iw x 3 BGR bytes
//   VideoWriter                 the same container
//   line()                      Bresenham, 1 px; every call is also appended to the file named by $RD_STUB_LOG as
//                               "line <x0> <y0> <x1> <y1> <b> <g> <r> <thickness>" (how the tests read the programs' results);
//                               VideoWriter::write appends "frame" to the same log
//   namedWindow / imshow / waitKey   no display: waitKey returns 27 (ESC)
#ifndef RD_OPENCV_STUB_HPP
#define RD_OPENCV_STUB_HPP
#include <assert.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <memory>
#include <string>
#include <vector>

#define CV_LOAD_IMAGE_COLOR 1
#define CV_CAP_PROP_FRAME_WIDTH 3
#define CV_CAP_PROP_FRAME_HEIGHT 4

struct CvPoint { int x, y; };
struct CvSize { int width, height; };
static inline CvPoint cvPoint(int x, int y) { CvPoint p = {x, y}; return p; }
static inline CvSize cvSize(int w, int h) { CvSize s = {w, h}; return s; }

namespace cv {
enum { WINDOW_AUTOSIZE = 1 };
struct Scalar {
  double val[4];
  Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
};
class Mat {
 public:
  uint8_t *data;
  int cols, rows;
  size_t step;
  Mat() : data(NULL), cols(0), rows(0), step(0) {}
  void create(int r, int c) {
    rows = r; cols = c; step = ((size_t)c * 3 + 3) & ~(size_t)3;
    store.reset(new std::vector<uint8_t>(step * (size_t)r, 0));
    data = store->data();
  }
  int channels() const { return 3; }
  Mat clone() const { Mat m; if (data) { m.create(rows, cols); memcpy(m.data, data, step * (size_t)rows); } return m; }
  void copyTo(Mat &m) const { if (m.rows != rows || m.cols != cols || !m.data) m.create(rows, cols); memcpy(m.data, data, step * (size_t)rows); }
 private:
  std::shared_ptr<std::vector<uint8_t> > store;
};
static inline void rd_stub_log(const char *fmt, int a = 0, int b = 0, int c = 0, int d = 0, int e = 0, int f = 0, int g = 0, int h = 0) {
  const char *path = getenv("RD_STUB_LOG");
  if (!path) return;
  FILE *fp = fopen(path, "a");
  if (!fp) return;
  fprintf(fp, fmt, a, b, c, d, e, f, g, h);
  fclose(fp);
}
static inline Mat imread(const std::string &path, int) {
  Mat m;
  FILE *fp = fopen(path.c_str(), "rb");
  if (!fp) return m;
  int w = 0, h = 0, mx = 0;
  if (fscanf(fp, "P6 %d %d %d", &w, &h, &mx) == 3 && mx == 255 && fgetc(fp) != EOF) {
    m.create(h, w);
    std::vector<uint8_t> row((size_t)w * 3);
    for (int y = 0; y < h; y++) {
      if (fread(row.data(), 1, row.size(), fp) != row.size()) { m = Mat(); break; }
      for (int x = 0; x < w; x++) { uint8_t *p = m.data + y * m.step + x * 3; p[0] = row[x * 3 + 2]; p[1] = row[x * 3 + 1]; p[2] = row[x * 3]; }
    }
  }
  fclose(fp);
  return m;
}
static inline bool imwrite(const std::string &path, const Mat &m) {
  FILE *fp = fopen(path.c_str(), "wb");
  if (!fp) return false;
  fprintf(fp, "P6\n%d %d\n255\n", m.cols, m.rows);
  for (int y = 0; y < m.rows; y++)
    for (int x = 0; x < m.cols; x++) { const uint8_t *p = m.data + y * m.step + x * 3; fputc(p[2], fp); fputc(p[1], fp); fputc(p[0], fp); }
  fclose(fp);
  return true;
}
static inline void line(Mat &img, CvPoint a, CvPoint b, const Scalar &c, int thickness = 1, int = 8, int = 0) {
  rd_stub_log("line %d %d %d %d %d %d %d %d\n", a.x, a.y, b.x, b.y, (int)c.val[0], (int)c.val[1], (int)c.val[2], thickness);
  int x = a.x, y = a.y;
  const int dx = abs(b.x - a.x), dy = -abs(b.y - a.y), sx = a.x < b.x ? 1 : -1, sy = a.y < b.y ? 1 : -1;
  int err = dx + dy;
  for (long guard = 0; guard < 100000; guard++) {
    if (x >= 0 && x < img.cols && y >= 0 && y < img.rows) { uint8_t *p = img.data + y * img.step + x * 3; p[0] = (uint8_t)c.val[0]; p[1] = (uint8_t)c.val[1]; p[2] = (uint8_t)c.val[2]; }
    if (x == b.x && y == b.y) break;
    const int e2 = 2 * err;
    if (e2 >= dy) { err += dy; x += sx; }
    if (e2 <= dx) { err += dx; y += sy; }
  }
}
class VideoCapture {
 public:
  explicit VideoCapture(int) : fp(NULL), iw(0), ih(0), n(0), have(false) {}          // no cameras here
  explicit VideoCapture(const std::string &path) : fp(fopen(path.c_str(), "rb")), iw(0), ih(0), n(0), have(false) {
    if (fp && (fscanf(fp, "RDV1 %d %d %d", &iw, &ih, &n) != 3 || fgetc(fp) == EOF)) { fclose(fp); fp = NULL; }
  }
  ~VideoCapture() { if (fp) fclose(fp); }
  bool isOpened() const { return fp != NULL; }
  double get(int prop) const { return prop == CV_CAP_PROP_FRAME_WIDTH ? iw : prop == CV_CAP_PROP_FRAME_HEIGHT ? ih : 0; }
  bool set(int, double) { return false; }
  bool grab() {
    have = false;
    if (!fp || n <= 0) return false;
    frame.create(ih, iw);
    for (int y = 0; y < ih; y++) if (fread(frame.data + y * frame.step, 1, (size_t)iw * 3, fp) != (size_t)iw * 3) return false;
    n--;
    have = true;
    return true;
  }
  bool retrieve(Mat &m, int = 0) { if (!have) return false; frame.copyTo(m); return true; }
 private:
  FILE *fp;
  int iw, ih, n;
  bool have;
  Mat frame;
};
class VideoWriter {
 public:
  VideoWriter(const std::string &path, int, double, CvSize s, bool) : fp(fopen(path.c_str(), "wb")) { if (fp) fprintf(fp, "RDV1 %d %d %d\n", s.width, s.height, 0); }
  ~VideoWriter() { if (fp) fclose(fp); }
  bool isOpened() const { return fp != NULL; }
  void write(const Mat &m) {
    rd_stub_log("frame\n");
    for (int y = 0; y < m.rows; y++) fwrite(m.data + y * m.step, 1, (size_t)m.cols * 3, fp);
    fflush(fp);
  }
 private:
  FILE *fp;
};
static inline void namedWindow(const std::string &, int = 0) {}
static inline void imshow(const std::string &, const Mat &) {}
static inline int waitKey(int = 0) { return 27; }
static inline void destroyAllWindows() {}
}  // namespace cv
#endif
