// Stand-in for <opencv2/core/eigen.hpp> (see ../opencv.hpp): element-wise copies between cv::Mat / cv::Matx and Eigen.  TEST INFRASTRUCTURE.
#pragma once
#include "../opencv.hpp"
#include <Eigen/Dense>
namespace cv {
template <class T, int R, int C, int O, int MR, int MC>
inline void cv2eigen(const Mat& src, Eigen::Matrix<T, R, C, O, MR, MC>& dst) {
    dst.resize(src.rows, src.cols);
    for (int i = 0; i < src.rows; ++i) for (int j = 0; j < src.cols; ++j) dst(i, j) = (T)src.at<double>(i, j);
}
template <class T, int M, int N, int R, int C, int O, int MR, int MC>
inline void cv2eigen(const Matx<T, M, N>& src, Eigen::Matrix<T, R, C, O, MR, MC>& dst) {
    dst.resize(M, N);
    for (int i = 0; i < M; ++i) for (int j = 0; j < N; ++j) dst(i, j) = src(i, j);
}
template <class T, int R, int C, int O, int MR, int MC>
inline void eigen2cv(const Eigen::Matrix<T, R, C, O, MR, MC>& src, Mat& dst) {
    dst.create((int)src.rows(), (int)src.cols(), DepthOf<T>::value);
    for (int i = 0; i < dst.rows; ++i) for (int j = 0; j < dst.cols; ++j) dst.at<T>(i, j) = src(i, j);
}
template <class T, int M, int N, int O, int MR, int MC>
inline void eigen2cv(const Eigen::Matrix<T, M, N, O, MR, MC>& src, Matx<T, M, N>& dst) {
    for (int i = 0; i < M; ++i) for (int j = 0; j < N; ++j) dst(i, j) = src(i, j);
}
}  // namespace cv
