// TEST INFRASTRUCTURE — cv::FileStorage / cv::FileNode as a closed door: the vendored DBoW2 declares virtual YAML
// save()/load() that must compile; nothing in the pinning tests opens a YAML vocabulary (they use the text format).
#ifndef B2S_REFSHIM_FILESTORAGE_H
#define B2S_REFSHIM_FILESTORAGE_H
#include <iostream>
#include <sstream>
#include <string>
namespace cv {
class FileNode {
 public:
  FileNode operator[](const std::string&) const { return FileNode(); }
  FileNode operator[](const char*) const { return FileNode(); }
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
  FileStorage(const std::string&, int) {}
  bool isOpened() const { return false; }
  void release() {}
  FileNode operator[](const std::string&) const { return FileNode(); }
  FileNode operator[](const char*) const { return FileNode(); }
};
template <typename T>
inline FileStorage& operator<<(FileStorage& fs, const T&) {
  return fs;
}
}  // namespace cv
#endif
