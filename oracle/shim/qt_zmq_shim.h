/* TEST INFRASTRUCTURE ONLY (oracle/): minimal stand-in for the few Qt symbols the
 * reference's publish/ hot-path sources touch, so that the UNMODIFIED files under
 * /root/reference/publish compile with plain g++.
 * Nothing here is product code; nothing here is copied from Qt. */
#ifndef AERODDC_ORACLE_QT_SHIM_H
#define AERODDC_ORACLE_QT_SHIM_H
#include <algorithm>
#include <cassert>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#define Q_OBJECT
#define signals public
#define slots
#define emit

class QObject {
public:
  explicit QObject(QObject *parent = nullptr) { (void)parent; }
  virtual ~QObject() {}
};

template <class T> class QVector : public std::vector<T> {
public:
  using std::vector<T>::vector;
  QVector() : std::vector<T>() {}
  int length() const { return (int)this->size(); }
};

class QByteArrayShim {
public:
  explicit QByteArrayShim(const std::string &s) : s_(s) {}
  const char *constData() const { return s_.c_str(); }
private:
  std::string s_;
};

class QString {
public:
  QString() {}
  QString(const char *c) : s_(c ? c : "") {}
  QString(const std::string &s) : s_(s) {}
  int compare(const QString &o) const { return s_.compare(o.s_); }
  int length() const { return (int)s_.size(); }
  QByteArrayShim toUtf8() const { return QByteArrayShim(s_); }
  bool operator==(const QString &o) const { return s_ == o.s_; }
  bool operator!=(const QString &o) const { return s_ != o.s_; }
private:
  std::string s_;
};
#endif
