// oracle/refshim/refshim_cv.h — TEST INFRASTRUCTURE ONLY.
// A minimal stand-in for the OpenCV C++ API used by /root/reference/src/ORBextractor.cc, so that the reference's OWN
// source file can be compiled here (OpenCV C++ headers are not installed) and its output compared with the oracle's
// restatement (tests/test_oracle_reference_extractor.py).  The OpenCV-owned arithmetic (resize, FAST, GaussianBlur,
// fastAtan2) is delegated to the oracle primitives that tests/test_oracle_cv2.py pins bit-for-bit against cv2 4.13; the
// containers (Mat with ROI views, KeyPoint, Point, ...) only reproduce OpenCV's semantics, not its code.
#pragma once
#include <algorithm>
#include <climits>
#include <cassert>
#include <cstdlib>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

#include "../orb_oracle.h"

typedef unsigned char uchar;
#define CV_PI 3.1415926535897932384626433832795
#define CV_8U 0
#define CV_8UC1 0
#define CV_32F 5
#define CV_32FC1 5

inline int cvRound(double v) { return (int)std::nearbyint(v); }  // round half to even (default rounding mode)
inline int cvFloor(double v) { return (int)std::floor(v); }
inline int cvCeil(double v) { return (int)std::ceil(v); }

namespace cv {
enum { INTER_LINEAR = 1 };
enum { BORDER_REFLECT_101 = 4, BORDER_ISOLATED = 16 };

template <typename T>
struct Point_ {
  T x, y;
  Point_() : x(0), y(0) {}
  Point_(T _x, T _y) : x(_x), y(_y) {}
  Point_& operator*=(float s) {
    x = (T)(x * s);
    y = (T)(y * s);
    return *this;
  }
};
typedef Point_<int> Point2i;
typedef Point2i Point;
typedef Point_<float> Point2f;
template <typename T>
struct Point3_ {
  T x, y, z;
  Point3_() : x(0), y(0), z(0) {}
  Point3_(T _x, T _y, T _z) : x(_x), y(_y), z(_z) {}
};
typedef Point3_<float> Point3f;

struct Size {
  int width, height;
  Size() : width(0), height(0) {}
  Size(int w, int h) : width(w), height(h) {}
};
struct Rect {
  int x, y, width, height;
  Rect(int _x, int _y, int w, int h) : x(_x), y(_y), width(w), height(h) {}
};

struct KeyPoint {  // 28-byte layout of cv::KeyPoint
  Point2f pt;
  float size, angle, response;
  int octave, class_id;
  KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {}
  KeyPoint(float x, float y, float _size, float _angle = -1, float _response = 0, int _octave = 0, int _class_id = -1)
      : pt(x, y), size(_size), angle(_angle), response(_response), octave(_octave), class_id(_class_id) {}
};

class Mat {
 public:
  struct Step {
    size_t v = 0;
    operator size_t() const { return v; }
  };
  int rows = 0, cols = 0;
  uchar* data = nullptr;
  Step step;  // bytes per row
  int type_ = CV_8UC1;
  size_t elemSize() const { return type_ == CV_32F ? 4 : 1; }

  Mat() {}
  Mat(int r, int c, int type) { create(r, c, type); }
  Mat(Size s, int type) { create(s.height, s.width, type); }
  // user-data constructor (no copy), as cv::Mat(rows, cols, type, data, step)
  Mat(int r, int c, int, void* d, size_t st) : rows(r), cols(c), data((uchar*)d) { step.v = st; }
  void create(int r, int c, int type) {
    if (r == rows && c == cols && data && type == type_) return;
    type_ = type;
    buf_.reset(new std::vector<uchar>((size_t)r * c * elemSize()));
    rows = r;
    cols = c;
    data = buf_->data();
    step.v = (size_t)c * elemSize();
  }
  void release() {
    buf_.reset();
    rows = cols = 0;
    data = nullptr;
    step.v = 0;
  }
  // cv::Mat::zeros returns a MatExpr: assigning it to a Mat of the same size FILLS THE EXISTING BUFFER (Mat::create does
  // not reallocate) — computeDescriptors (src/ORBextractor.cc:1531) relies on that to write into a rowRange view
  struct ZerosExpr {
    int rows, cols, type;
  };
  static ZerosExpr zeros(int r, int c, int type) { return ZerosExpr{r, c, type}; }
  Mat(const ZerosExpr& z) { *this = z; }
  Mat& operator=(const ZerosExpr& z) {
    if (!(z.rows == rows && z.cols == cols && data)) {
      release();
      create(z.rows, z.cols, z.type);
    }
    for (int r = 0; r < rows; r++) std::memset(data + (size_t)r * step.v, 0, cols * elemSize());
    return *this;
  }
  bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
  int type() const { return type_; }
  size_t step1() const { return step.v / elemSize(); }
  Mat operator()(const Rect& r) const {
    Mat m = *this;  // shares the buffer
    m.data = data + (size_t)r.y * step.v + (size_t)r.x * elemSize();
    m.rows = r.height;
    m.cols = r.width;
    return m;
  }
  Mat rowRange(int a, int b) const { return (*this)(Rect(0, a, cols, b - a)); }
  Mat colRange(int a, int b) const { return (*this)(Rect(a, 0, b - a, rows)); }
  Mat clone() const {
    Mat m(rows, cols, type_);
    for (int r = 0; r < rows; r++) std::memcpy(m.data + (size_t)r * m.step.v, data + (size_t)r * step.v, cols * elemSize());
    return m;
  }
  Mat reshape(int) const { return *this; }  // only reached on the distortion path, which the pinning tests do not take
  static Mat eye(int r, int c, int type) {  // CV_32F only (include/Converter.h users)
    Mat m(r, c, type);
    for (int i = 0; i < r; i++)
      for (int j = 0; j < c; j++) m.at<float>(i, j) = i == j ? 1.0f : 0.0f;
    return m;
  }
  static Mat ones(int r, int c, int type) {
    Mat m(r, c, type);
    for (int i = 0; i < r; i++)
      for (int j = 0; j < c; j++) m.at<float>(i, j) = 1.0f;
    return m;
  }
  // u8 -> f32 (the only conversion the reference asks for); a differently-typed destination is reallocated, so
  // converting a pyramid ROI "in place" leaves the pyramid untouched, as in OpenCV
  void convertTo(Mat& dst, int type) const {
    Mat m(rows, cols, type);
    for (int r = 0; r < rows; r++)
      for (int c = 0; c < cols; c++) {
        const float v = type_ == CV_32F ? at<float>(r, c) : (float)at<uchar>(r, c);
        if (type == CV_32F)
          m.at<float>(r, c) = v;
        else
          m.at<uchar>(r, c) = (uchar)v;
      }
    dst = m;
  }
  void copyTo(Mat& dst) const {  // deep copy into a (re)allocated destination
    Mat m = clone();
    dst = m;
  }
  Mat row(int r) const { return rowRange(r, r + 1); }
  Mat col(int c) const { return colRange(c, c + 1); }
  template <typename T>
  T* ptr(int r = 0) {
    return (T*)(data + (size_t)r * step.v);
  }
  template <typename T>
  const T* ptr(int r = 0) const {
    return (const T*)(data + (size_t)r * step.v);
  }
  // single-index access of a vector (n x 1 or 1 x n), as cv::Mat::at<T>(int)
  template <typename T>
  T& at(int i) {
    return rows == 1 ? at<T>(0, i) : at<T>(i, 0);
  }
  template <typename T>
  const T& at(int i) const {
    return rows == 1 ? at<T>(0, i) : at<T>(i, 0);
  }
  float f(int r, int c) const { return at<float>(r, c); }
  Mat t() const {
    Mat m(cols, rows, type_);
    for (int r = 0; r < rows; r++)
      for (int c = 0; c < cols; c++) m.at<float>(c, r) = at<float>(r, c);
    return m;
  }
  double dot(const Mat& b) const {  // cv::Mat::dot: double accumulation of the float products
    double s = 0;
    for (int r = 0; r < rows; r++)
      for (int c = 0; c < cols; c++) s += (double)at<float>(r, c) * (double)b.at<float>(r, c);
    return s;
  }
  template <typename T>
  T& at(int r, int c) {
    return *(T*)(data + (size_t)r * step.v + (size_t)c * sizeof(T));
  }
  template <typename T>
  const T& at(int r, int c) const {
    return *(const T*)(data + (size_t)r * step.v + (size_t)c * sizeof(T));
  }
  uchar* ptr(int r = 0) { return data + (size_t)r * step.v; }
  const uchar* ptr(int r = 0) const { return data + (size_t)r * step.v; }

 private:
  std::shared_ptr<std::vector<uchar> > buf_;
};

// ---- CV_32F matrix expressions as OpenCV evaluates them: gemm with double accumulators, scaling by a double alpha
inline Mat operator*(const Mat& a, const Mat& b) {
  Mat m(a.rows, b.cols, CV_32F);
  for (int r = 0; r < a.rows; r++)
    for (int c = 0; c < b.cols; c++) {
      double s = 0;
      for (int k = 0; k < a.cols; k++) s += (double)a.f(r, k) * (double)b.f(k, c);
      m.at<float>(r, c) = (float)s;
    }
  return m;
}
inline Mat scaled(const Mat& a, double alpha) {
  Mat m(a.rows, a.cols, CV_32F);
  for (int r = 0; r < a.rows; r++)
    for (int c = 0; c < a.cols; c++) m.at<float>(r, c) = (float)((double)a.f(r, c) * alpha);
  return m;
}
inline Mat operator*(double s, const Mat& a) { return scaled(a, s); }
inline Mat operator*(const Mat& a, double s) { return scaled(a, s); }
inline Mat operator/(const Mat& a, double s) { return scaled(a, 1.0 / s); }
inline Mat operator-(const Mat& a) { return scaled(a, -1.0); }
inline Mat operator+(const Mat& a, const Mat& b) {
  Mat m(a.rows, a.cols, CV_32F);
  for (int r = 0; r < a.rows; r++)
    for (int c = 0; c < a.cols; c++) m.at<float>(r, c) = a.f(r, c) + b.f(r, c);
  return m;
}
inline Mat operator-(const Mat& a, const Mat& b) {
  Mat m(a.rows, a.cols, CV_32F);
  for (int r = 0; r < a.rows; r++)
    for (int c = 0; c < a.cols; c++) m.at<float>(r, c) = a.f(r, c) - b.f(r, c);
  return m;
}
inline double norm(const Mat& a) {  // NORM_L2
  double s = 0;
  for (int r = 0; r < a.rows; r++)
    for (int c = 0; c < a.cols; c++) s += (double)a.f(r, c) * (double)a.f(r, c);
  return std::sqrt(s);
}

enum { NORM_L1 = 2, NORM_L2 = 4 };
inline double norm(const Mat& a, const Mat& b, int type) {  // NORM_L1 of a CV_32F difference, double accumulator
  double s = 0;
  for (int r = 0; r < a.rows; r++)
    for (int c = 0; c < a.cols; c++) {
      const double d = (double)a.f(r, c) - (double)b.f(r, c);
      s += type == NORM_L1 ? std::fabs(d) : d * d;
    }
  return type == NORM_L1 ? s : std::sqrt(s);
}
// cv::Mat_<float>(r, c) << a, b, c
template <typename T>
class Mat_ : public Mat {
 public:
  Mat_(int r, int c) : Mat(r, c, CV_32F) {}
  struct Comma {
    Mat m;
    int i;
    Comma operator,(T v) {
      m.at<T>(i / m.cols, i % m.cols) = v;
      return Comma{m, i + 1};
    }
    operator Mat() const { return m; }
  };
  Comma operator<<(T v) {
    at<T>(0, 0) = v;
    return Comma{*this, 1};
  }
};
class _InputArray {
 public:
  _InputArray() : m_(nullptr) {}
  _InputArray(const Mat& m) : m_(const_cast<Mat*>(&m)) {}
  bool empty() const { return !m_ || m_->empty(); }
  Mat getMat() const { return m_ ? *m_ : Mat(); }

 protected:
  Mat* m_;
};
class _OutputArray : public _InputArray {
 public:
  _OutputArray(Mat& m) { m_ = &m; }
  void create(int r, int c, int type) const { m_->create(r, c, type); }
  void release() const { m_->release(); }
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;

inline void undistortPoints(const Mat&, Mat&, const Mat&, const Mat&, const Mat&, const Mat&) {
  std::abort();  // Frame::UndistortKeyPoints returns before this when mDistCoef[0] == 0 (src/Frame.cc:905-909)
}
inline float fastAtan2(float y, float x) { return orc_fast_atan2(y, x); }

inline void resize(const Mat& src, Mat& dst, Size dsize, double, double, int interpolation) {
  assert(interpolation == INTER_LINEAR && dst.rows == dsize.height && dst.cols == dsize.width);
  (void)interpolation;
  orc_resize_linear_u8(src.data, src.cols, src.rows, (int)src.step.v, dst.data, dsize.width, dsize.height, (int)dst.step.v);
}

// cv::copyMakeBorder with BORDER_REFLECT_101 (+BORDER_ISOLATED): dst is (rows+top+bottom) x (cols+left+right); src may be
// the interior ROI of dst itself
inline void copyMakeBorder(const Mat& src, Mat& dst, int top, int bottom, int left, int right, int) {
  const int R = src.rows, C = src.cols;
  if (dst.rows != R + top + bottom || dst.cols != C + left + right) dst.create(R + top + bottom, C + left + right, CV_8UC1);
  uchar* inner = dst.data + (size_t)top * dst.step.v + left;
  if (inner != src.data)
    for (int r = 0; r < R; r++) std::memmove(inner + (size_t)r * dst.step.v, src.data + (size_t)r * src.step.v, C);
  auto refl = [](int p, int n) {
    if (n == 1) return 0;
    while (p < 0 || p >= n) p = (p < 0) ? -p : 2 * (n - 1) - p;
    return p;
  };
  for (int r = 0; r < R; r++) {
    uchar* row = inner + (size_t)r * dst.step.v;
    for (int c = -left; c < 0; c++) row[c] = row[refl(c, C)];
    for (int c = C; c < C + right; c++) row[c] = row[refl(c, C)];
  }
  const size_t W = (size_t)C + left + right;
  for (int r = -top; r < 0; r++)
    std::memcpy(dst.data + (size_t)(r + top) * dst.step.v, dst.data + (size_t)(refl(r, R) + top) * dst.step.v, W);
  for (int r = R; r < R + bottom; r++)
    std::memcpy(dst.data + (size_t)(r + top) * dst.step.v, dst.data + (size_t)(refl(r, R) + top) * dst.step.v, W);
}

inline void GaussianBlur(const Mat& src, Mat& dst, Size ksize, double sx, double sy, int) {
  assert(ksize.width == 7 && ksize.height == 7 && sx == 2 && sy == 2);
  (void)ksize; (void)sx; (void)sy;
  Mat out(src.rows, src.cols, CV_8UC1);
  orc_gaussian_blur7_s2_u8(src.data, src.cols, src.rows, (int)src.step.v, out.data, (int)out.step.v);
  if (dst.rows != src.rows || dst.cols != src.cols || !dst.data) dst.create(src.rows, src.cols, CV_8UC1);
  for (int r = 0; r < src.rows; r++) std::memcpy(dst.data + (size_t)r * dst.step.v, out.data + (size_t)r * out.step.v, src.cols);
}

// cv::FAST(img, keypoints, threshold, nonmaxSuppression=true): KeyPoint(x, y, 7.f, -1, score) in row-major order
inline void FAST(const Mat& img, std::vector<KeyPoint>& keypoints, int threshold, bool nonmax = true) {
  assert(nonmax);
  (void)nonmax;
  keypoints.clear();
  const int cap = img.rows * img.cols + 1;
  std::vector<int> xy((size_t)cap * 2), resp(cap);
  const int n = orc_fast9_16_nms(img.data, img.cols, img.rows, (int)img.step.v, threshold, xy.data(), resp.data(), cap);
  keypoints.reserve(n);
  for (int i = 0; i < n; i++) keypoints.push_back(KeyPoint((float)xy[2 * i], (float)xy[2 * i + 1], 7.f, -1, (float)resp[i]));
}
// only referenced by ORBextractor::ComputeKeyPointsOld, which operator() never calls
struct KeyPointsFilter {
  static void retainBest(std::vector<KeyPoint>&, int) { std::abort(); }
};
}  // namespace cv
