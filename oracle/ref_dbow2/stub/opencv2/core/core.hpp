// Minimal stand-in for <opencv2/core/core.hpp>, ONLY so that the reference's vendored DBoW2 sources
// (/root/reference/ThirdParty/DBoW2) compile in a container without OpenCV C++ headers.  TEST INFRASTRUCTURE.
// It provides just what those sources touch: a byte-matrix cv::Mat (create/zeros/ptr/clone/release/row) and
// inert FileStorage/FileNode types (the YAML load/save paths are compiled but never called: the vocabulary is
// loaded through the reference's own loadFromTextFile).
#pragma once
#include <algorithm>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <iostream>
#include <limits>
#include <sstream>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#define CV_8U 0
#define CV_32F 5

namespace cv {

class Mat {
public:
    int rows = 0, cols = 0;
    Mat() {}
    Mat(int r, int c, int type) { create(r, c, type); }
    void create(int r, int c, int type) {
        esz_ = type == CV_32F ? 4 : 1;
        rows = r; cols = c;
        buf_ = std::shared_ptr<uint8_t>(new uint8_t[(size_t)r * c * esz_ + 8](), std::default_delete<uint8_t[]>());
        data_ = buf_.get();
    }
    static Mat zeros(int r, int c, int type) { return Mat(r, c, type); }
    void release() { buf_.reset(); data_ = nullptr; rows = cols = 0; }
    bool empty() const { return data_ == nullptr || rows == 0 || cols == 0; }
    Mat clone() const {
        Mat m;
        if (!empty()) { m.create(rows, cols, esz_ == 4 ? CV_32F : CV_8U); std::memcpy(m.data_, data_, (size_t)rows * cols * esz_); }
        return m;
    }
    Mat row(int r) const { Mat m; m.buf_ = buf_; m.esz_ = esz_; m.rows = 1; m.cols = cols; m.data_ = data_ + (size_t)r * cols * esz_; return m; }
    template <class T> T* ptr(int r = 0) { return (T*)(data_ + (size_t)r * cols * esz_); }
    template <class T> const T* ptr(int r = 0) const { return (const T*)(data_ + (size_t)r * cols * esz_); }
private:
    std::shared_ptr<uint8_t> buf_;
    uint8_t* data_ = nullptr;
    int esz_ = 1;
};

class FileNode {
public:
    FileNode operator[](const char*) const { return FileNode(); }
    FileNode operator[](const std::string&) const { return FileNode(); }
    FileNode operator[](int) const { return FileNode(); }
    size_t size() const { return 0; }
    operator int() const { return 0; }
    operator float() const { return 0.f; }
    operator double() const { return 0.0; }
    operator std::string() const { return std::string(); }
};

class FileStorage {
public:
    enum { READ = 0, WRITE = 1 };
    FileStorage() {}
    FileStorage(const char*, int) { std::abort(); }          // never reached by the oracle harness
    FileStorage(const std::string&, int) { std::abort(); }
    bool isOpened() const { return false; }
    FileNode operator[](const char*) const { return FileNode(); }
    FileNode operator[](const std::string&) const { return FileNode(); }
};
template <class T> inline FileStorage& operator<<(FileStorage& fs, const T&) { return fs; }

}  // namespace cv
