// stand-in for <opencv2/core/mat.hpp>: declarations only (see README.md)
#pragma once
#include <cstddef>
#define CV_32F 5
#define CV_64F 6
namespace cv {
struct Rect { Rect(int, int, int, int); };
class Mat {
   public:
    Mat();
    Mat clone() const;
    void convertTo(Mat& m, int rtype) const;
    template <class T> T& at(int i, int j);
    template <class T> T& at(int i0 = 0);
    size_t total() const;
    Mat operator()(const Rect& roi) const;
    Mat inv() const;
    bool isContinuous() const;
    unsigned char* data;
    int rows, cols;
};
Mat operator*(const Mat& a, const Mat& b);
}  // namespace cv
