// TEST INFRASTRUCTURE (oracle/): the extra Qt pieces /root/reference/publish/publisher.{h,cpp} needs on top of
// oracle/shim_decode/qt_decode_shim.h - QSettings (IniFormat reader), QFileInfo, QtConcurrent::run, QFuture,
// QSocketNotifier and printf-style qDebug - so that the reference's Publisher (settings-file semantics,
// DC correction, main -> sub VFO tree) compiles UNMODIFIED into oracle/_ref/ref_publish. Written from the Qt
// documentation; nothing here is reference code.
#ifndef AERODDC_QT_PUBLISHER_SHIM_H
#define AERODDC_QT_PUBLISHER_SHIM_H

#include <sys/stat.h>

#include <fstream>
#include <map>
#include <sstream>

#include "../shim_decode/qt_decode_shim.h"

// common/logger.h logs with qDebug("format", args...)
inline void qDebug(const char* fmt, ...) __attribute__((format(printf, 1, 2)));
inline void qDebug(const char* fmt, ...) {
  if (!getenv("REF_PUBLISH_VERBOSE")) return;
  va_list ap;
  va_start(ap, fmt);
  vfprintf(stderr, fmt, ap);
  va_end(ap);
  fputc('\n', stderr);
}

class QVariant {
public:
  QVariant() : valid_(false) {}
  explicit QVariant(const std::string& s) : s_(s), valid_(true) {}
  // QVariant(QString).toInt(): the whole string must be an integer, else 0
  int toInt(bool* ok = nullptr) const {
    char* e = nullptr;
    const std::string t = QByteArray(s_).trimmed().d;
    const long long v = std::strtoll(t.c_str(), &e, 10);
    const bool good = valid_ && !t.empty() && e && *e == 0;
    if (ok) *ok = good;
    return good ? (int)v : 0;
  }
  float toFloat(bool* ok = nullptr) const {
    char* e = nullptr;
    const std::string t = QByteArray(s_).trimmed().d;
    const float v = std::strtof(t.c_str(), &e);
    const bool good = valid_ && !t.empty() && e && *e == 0;
    if (ok) *ok = good;
    return good ? v : 0.0f;
  }
  double toDouble(bool* ok = nullptr) const { return (double)toFloat(ok); }
  QString toString() const { return QString(s_); }
  bool isValid() const { return valid_; }
private:
  std::string s_;
  bool valid_;
};

// QSettings, IniFormat, read-only: "[General]" (or no section) is the top level, "[group]" prefixes its keys with
// "group/", a backslash in a key is the path separator (arrays are stored as "name/size" and "name/<1-based index>/key"),
// ';' starts a comment line, values may be double-quoted.
class QSettings {
public:
  enum Format { NativeFormat, IniFormat };
  QSettings(const QString& path, Format) {
    std::ifstream f(path.toStdString());
    std::string line, group;
    while (std::getline(f, line)) {
      line = QByteArray(line).trimmed().d;
      if (line.empty() || line[0] == ';' || line[0] == '#') continue;
      if (line[0] == '[') {
        const size_t e = line.find(']');
        group = QByteArray(line.substr(1, e == std::string::npos ? std::string::npos : e - 1)).trimmed().d;
        if (group == "General") group.clear();
        continue;
      }
      const size_t eq = line.find('=');
      if (eq == std::string::npos) continue;
      std::string key = QByteArray(line.substr(0, eq)).trimmed().d, val = QByteArray(line.substr(eq + 1)).trimmed().d;
      if (val.size() >= 2 && val.front() == '"' && val.back() == '"') val = val.substr(1, val.size() - 2);
      for (char& c : key)
        if (c == '\\') c = '/';
      kv_[group.empty() ? key : group + "/" + key] = val;
    }
  }
  QVariant value(const QString& key) const {
    auto it = kv_.find(scoped(key.toStdString()));
    return it == kv_.end() ? QVariant() : QVariant(it->second);
  }
  bool contains(const QString& key) const { return kv_.count(scoped(key.toStdString())) != 0; }
  int beginReadArray(const QString& name) {
    array_ = name.toStdString();
    index_ = -1;
    return value("size").toInt();
  }
  void setArrayIndex(int i) { index_ = i; }
  void endArray() { array_.clear(); index_ = -1; }
private:
  std::string scoped(const std::string& key) const {
    if (array_.empty()) return key;
    if (index_ < 0) return array_ + "/" + key;
    return array_ + "/" + std::to_string(index_ + 1) + "/" + key;
  }
  std::map<std::string, std::string> kv_;
  std::string array_;
  int index_ = -1;
};

class QFileInfo {
public:
  explicit QFileInfo(const QString& path) : ok_(::stat(path.toStdString().c_str(), &st_) == 0) {}
  bool exists() const { return ok_; }
  bool isFile() const { return ok_ && S_ISREG(st_.st_mode); }
private:
  struct stat st_;
  bool ok_;
};

class QSocketNotifier : public QObject {};

template <class T> class QFuture {
public:
  void waitForFinished() {}
  bool isRunning() const { return false; }
  bool isFinished() const { return true; }
};
namespace QtConcurrent {
// the harness is single-threaded: the reader "thread" runs to the end of the stream inside run()
template <class F> QFuture<void> run(F f) { f(); return QFuture<void>(); }
}

#endif
