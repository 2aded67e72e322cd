// TEST INFRASTRUCTURE (oracle/): a minimal stand-in for the Qt types the reference's aero-decode sources use, so
// that decode/{aerol,mskdemodulator,oqpskdemodulator,DSP,coarsefreqestimate,jconvolutionalcodec,jfft,fftwrapper,
// hunter,databasetext}.cpp compile UNMODIFIED from /root/reference into oracle/_ref/libref_decode.so (SURVEY.md
// section 8d "end-to-end": GPU payloads into the unchanged decoder). Written from the Qt API documentation as the
// sources use it; nothing here is reference code. Strings are Latin-1 byte strings; signals are ordinary member
// functions whose bodies (what moc would generate) live in oracle/ref_decode_harness.cpp; connect() is a no-op
// because the harness routes every signal explicitly; timers never fire; QElapsedTimer reads 0 (only GUI
// refresh throttles look at it), which also makes the decoder deterministic.
#ifndef AERODDC_QT_DECODE_SHIM_H
#define AERODDC_QT_DECODE_SHIM_H

#include <algorithm>
#include <cassert>
#include <cctype>
#include <cmath>
#include <complex>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

typedef uint8_t quint8;
typedef uint16_t quint16;
typedef uint32_t quint32;
typedef uint64_t quint64;
typedef int8_t qint8;
typedef int16_t qint16;
typedef int32_t qint32;
typedef int64_t qint64;
typedef long long qlonglong;
typedef unsigned long long qulonglong;
typedef unsigned char uchar;
typedef unsigned short ushort;
typedef unsigned int uint;
typedef double qreal;

#define Q_OBJECT
#define Q_ENUM(x)
#define Q_UNUSED(x) (void)x;
#define signals public
#define slots
#define emit
#define SIGNAL(x) #x
#define SLOT(x) #x
#define foreach(var, container) for (var : container)

inline int qRound(double d) { return d >= 0.0 ? (int)(d + 0.5) : (int)(d - (double)((int)(d - 1)) + 0.5) + (int)(d - 1); }
template <class T> inline const T& qMin(const T& a, const T& b) { return a < b ? a : b; }
template <class T> inline const T& qMax(const T& a, const T& b) { return a < b ? b : a; }
template <class T> inline T qAbs(const T& a) { return a < 0 ? -a : a; }
template <class T> inline T qFromBigEndian(T v) {
  T r = 0;
  for (size_t i = 0; i < sizeof(T); i++) r |= ((v >> (8 * i)) & 0xFF) << (8 * (sizeof(T) - 1 - i));
  return r;
}
template <class T> inline T qToBigEndian(T v) { return qFromBigEndian(v); }
template <class T> inline T qFromLittleEndian(T v) { return v; }
template <class T> inline T qToLittleEndian(T v) { return v; }

class QString;
class QByteArray;

class QChar {
public:
  QChar() : c(0) {}
  QChar(char ch) : c((uchar)ch) {}
  QChar(int ch) : c((ushort)ch) {}
  QChar(uchar ch) : c(ch) {}
  char toLatin1() const { return c < 256 ? (char)c : 0; }
  ushort unicode() const { return c; }
  bool isPrint() const { return c < 256 && std::isprint(c); }
  bool isLetterOrNumber() const { return c < 256 && std::isalnum(c); }
  bool isDigit() const { return c < 256 && std::isdigit(c); }
  bool isSpace() const { return c < 256 && std::isspace(c); }
  bool operator==(QChar o) const { return c == o.c; }
  bool operator!=(QChar o) const { return c != o.c; }
  ushort c;
};
inline bool operator==(QChar a, char b) { return a.c == (uchar)b; }
inline bool operator!=(QChar a, char b) { return a.c != (uchar)b; }

template <class T> class QVector : public std::vector<T> {
public:
  typedef std::vector<T> base;
  using base::base;
  QVector() {}
  QVector(const base& b) : base(b) {}
  int size() const { return (int)base::size(); }
  int length() const { return (int)base::size(); }
  int count() const { return (int)base::size(); }
  bool isEmpty() const { return base::empty(); }
  const T& at(int i) const { return base::at((size_t)i); }
  T& operator[](int i) { assert(i >= 0 && i < size()); return base::operator[]((size_t)i); }
  const T& operator[](int i) const { assert(i >= 0 && i < size()); return base::operator[]((size_t)i); }
  void append(const T& v) { base::push_back(v); }
  void append(const QVector<T>& v) { base::insert(base::end(), v.begin(), v.end()); }
  void prepend(const T& v) { base::insert(base::begin(), v); }
  void push_front(const T& v) { base::insert(base::begin(), v); }
  void pop_front() { base::erase(base::begin()); }
  void removeFirst() { base::erase(base::begin()); }
  void removeLast() { base::pop_back(); }
  void removeAt(int i) { base::erase(base::begin() + i); }
  void remove(int i) { base::erase(base::begin() + i); }
  void remove(int i, int n) { base::erase(base::begin() + i, base::begin() + i + n); }
  void insert(int i, const T& v) { base::insert(base::begin() + i, v); }
  void replace(int i, const T& v) { (*this)[i] = v; }
  T takeFirst() { T v = base::front(); base::erase(base::begin()); return v; }
  T takeLast() { T v = base::back(); base::pop_back(); return v; }
  T takeAt(int i) { T v = (*this)[i]; removeAt(i); return v; }
  T& first() { return base::front(); }
  const T& first() const { return base::front(); }
  T& last() { return base::back(); }
  const T& last() const { return base::back(); }
  T value(int i) const { return (i >= 0 && i < size()) ? base::operator[]((size_t)i) : T(); }
  QVector<T>& fill(const T& v, int n = -1) {
    if (n >= 0) base::resize((size_t)n);
    std::fill(base::begin(), base::end(), v);
    return *this;
  }
  QVector<T> mid(int pos, int len = -1) const {
    if (pos < 0) pos = 0;
    if (pos > size()) pos = size();
    if (len < 0 || pos + len > size()) len = size() - pos;
    return QVector<T>(base(base::begin() + pos, base::begin() + pos + len));
  }
  int indexOf(const T& v, int from = 0) const {
    for (int i = from < 0 ? 0 : from; i < size(); i++)
      if (base::operator[]((size_t)i) == v) return i;
    return -1;
  }
  bool contains(const T& v) const { return indexOf(v) >= 0; }
  QVector<T>& operator<<(const T& v) { base::push_back(v); return *this; }
  QVector<T>& operator+=(const T& v) { base::push_back(v); return *this; }
  QVector<T>& operator+=(const QVector<T>& v) { append(v); return *this; }
  void squeeze() {}
  base toStdVector() const { return *this; }
};
// QVector<bool> would inherit std::vector<bool>'s proxy references; the sources do not use it.

template <class T> class QList : public QVector<T> {
public:
  using QVector<T>::QVector;
  QList() {}
  QList(const QVector<T>& v) : QVector<T>(v) {}
};

class QByteArray {
public:
  QByteArray() {}
  QByteArray(const char* s) : d(s ? s : "") {}
  QByteArray(const char* s, int n) : d(s, (size_t)(n < 0 ? std::strlen(s) : n)) {}
  QByteArray(int n, char c) : d((size_t)n, c) {}
  QByteArray(const std::string& s) : d(s) {}
  int size() const { return (int)d.size(); }
  int length() const { return (int)d.size(); }
  int count() const { return (int)d.size(); }
  bool isEmpty() const { return d.empty(); }
  void clear() { d.clear(); }
  void resize(int n) { d.resize((size_t)n); }
  void reserve(int n) { d.reserve((size_t)n); }
  char* data() { return d.empty() ? const_cast<char*>(d.c_str()) : &d[0]; }
  const char* data() const { return d.c_str(); }
  const char* constData() const { return d.c_str(); }
  operator const char*() const { return d.c_str(); }
  operator const void*() const { return d.c_str(); }
  char at(int i) const { assert(i >= 0 && i < size()); return d[(size_t)i]; }
  // Qt's non-const operator[] returns a QByteRef that grows the array on out-of-range assignment; the sources only
  // index inside the array.
  char& operator[](int i) { assert(i >= 0 && i < size()); return d[(size_t)i]; }
  char operator[](int i) const { assert(i >= 0 && i < size()); return d[(size_t)i]; }
  char& operator[](uint i) { assert(i < (uint)size()); return d[(size_t)i]; }
  char operator[](uint i) const { assert(i < (uint)size()); return d[(size_t)i]; }
  QByteArray& append(const QByteArray& o) { d += o.d; return *this; }
  QByteArray& append(char c) { d += c; return *this; }
  QByteArray& append(const char* s) { d += s; return *this; }
  QByteArray& append(const char* s, int n) { d.append(s, (size_t)n); return *this; }
  QByteArray& append(int n, char c) { d.append((size_t)n, c); return *this; }
  QByteArray& prepend(const QByteArray& o) { d = o.d + d; return *this; }
  QByteArray& prepend(char c) { d.insert(d.begin(), c); return *this; }
  void push_back(char c) { d += c; }
  void push_back(const QByteArray& o) { d += o.d; }
  void push_back(const char* s) { d += s; }
  void push_front(char c) { d.insert(d.begin(), c); }
  QByteArray& operator+=(const QByteArray& o) { d += o.d; return *this; }
  QByteArray& operator+=(char c) { d += c; return *this; }
  QByteArray& operator+=(const char* s) { d += s; return *this; }
  QByteArray& operator+=(const QString& s);
  QByteArray& fill(char c, int n = -1) {
    if (n >= 0) d.resize((size_t)n);
    std::fill(d.begin(), d.end(), c);
    return *this;
  }
  QByteArray mid(int pos, int len = -1) const {
    if (pos < 0) pos = 0;
    if (pos >= size()) return QByteArray();
    if (len < 0 || pos + len > size()) len = size() - pos;
    return QByteArray(d.substr((size_t)pos, (size_t)len));
  }
  QByteArray left(int n) const { return n >= size() ? *this : QByteArray(d.substr(0, (size_t)(n < 0 ? 0 : n))); }
  QByteArray right(int n) const { return n >= size() ? *this : QByteArray(d.substr(d.size() - (size_t)(n < 0 ? 0 : n))); }
  QByteArray& remove(int pos, int n) {
    if (pos >= 0 && pos < size()) d.erase((size_t)pos, (size_t)n);
    return *this;
  }
  QByteArray& insert(int pos, char c) { d.insert(d.begin() + pos, c); return *this; }
  void chop(int n) { d.resize(n >= size() ? 0 : d.size() - (size_t)n); }
  void truncate(int n) { if (n < size()) d.resize((size_t)(n < 0 ? 0 : n)); }
  QByteArray trimmed() const {
    size_t a = 0, b = d.size();
    while (a < b && std::isspace((uchar)d[a])) a++;
    while (b > a && std::isspace((uchar)d[b - 1])) b--;
    return QByteArray(d.substr(a, b - a));
  }
  bool startsWith(const QByteArray& o) const { return d.compare(0, o.d.size(), o.d) == 0; }
  bool endsWith(const QByteArray& o) const { return d.size() >= o.d.size() && d.compare(d.size() - o.d.size(), o.d.size(), o.d) == 0; }
  bool contains(char c) const { return d.find(c) != std::string::npos; }
  bool contains(const QByteArray& o) const { return d.find(o.d) != std::string::npos; }
  int indexOf(char c, int from = 0) const { size_t p = d.find(c, (size_t)from); return p == std::string::npos ? -1 : (int)p; }
  int indexOf(const QByteArray& o, int from = 0) const { size_t p = d.find(o.d, (size_t)from); return p == std::string::npos ? -1 : (int)p; }
  QByteArray toHex() const {
    static const char* h = "0123456789abcdef";
    std::string r;
    for (uchar c : d) { r += h[c >> 4]; r += h[c & 15]; }
    return QByteArray(r);
  }
  QByteArray toUpper() const { std::string r = d; for (auto& c : r) c = (char)std::toupper((uchar)c); return QByteArray(r); }
  QByteArray toLower() const { std::string r = d; for (auto& c : r) c = (char)std::tolower((uchar)c); return QByteArray(r); }
  int toInt(bool* ok = nullptr, int base = 10) const {
    char* e = nullptr;
    long v = std::strtol(d.c_str(), &e, base);
    if (ok) *ok = e && *e == 0 && !d.empty();
    return (int)v;
  }
  std::string toStdString() const { return d; }
  static QByteArray number(int v, int base = 10) {
    char b[40];
    std::snprintf(b, sizeof b, base == 16 ? "%x" : "%d", v);
    return QByteArray(b);
  }
  static QByteArray fromRawData(const char* s, int n) { return QByteArray(s, n); }
  std::string::iterator begin() { return d.begin(); }
  std::string::iterator end() { return d.end(); }
  std::string::const_iterator begin() const { return d.begin(); }
  std::string::const_iterator end() const { return d.end(); }
  bool operator==(const QByteArray& o) const { return d == o.d; }
  bool operator!=(const QByteArray& o) const { return d != o.d; }
  bool operator<(const QByteArray& o) const { return d < o.d; }
  bool operator==(const char* s) const { return d == s; }
  bool operator!=(const char* s) const { return d != s; }
  std::string d;
};
inline QByteArray operator+(const QByteArray& a, const QByteArray& b) { QByteArray r(a); r += b; return r; }
inline QByteArray operator+(const QByteArray& a, const char* b) { QByteArray r(a); r += b; return r; }
inline QByteArray operator+(const char* a, const QByteArray& b) { QByteArray r(a); r += b; return r; }
inline QByteArray operator+(const QByteArray& a, char b) { QByteArray r(a); r += b; return r; }

class QStringList;

class QString {
public:
  QString() {}
  QString(const char* s) : d(s ? s : "") {}
  QString(const std::string& s) : d(s) {}
  QString(const QByteArray& b) : d(b.d.c_str()) {}   // stops at the first NUL like QString::fromUtf8(const char*)
  QString(QChar c) : d(1, c.toLatin1()) {}
  QString(int n, QChar c) : d((size_t)n, c.toLatin1()) {}
  int size() const { return (int)d.size(); }
  int length() const { return (int)d.size(); }
  int count() const { return (int)d.size(); }
  bool isEmpty() const { return d.empty(); }
  bool isNull() const { return d.empty(); }
  void clear() { d.clear(); }
  void reserve(int n) { d.reserve((size_t)n); }
  void chop(int n) { d.resize(n >= size() ? 0 : d.size() - (size_t)n); }
  void truncate(int n) { if (n < size()) d.resize((size_t)(n < 0 ? 0 : n)); }
  QChar at(int i) const { return QChar(d[(size_t)i]); }
  QChar operator[](int i) const { return QChar(d[(size_t)i]); }
  char& operator[](int i) { return d[(size_t)i]; }
  QString& operator+=(const QString& o) { d += o.d; return *this; }
  QString& operator+=(const char* s) { d += s; return *this; }
  QString& operator+=(char c) { d += c; return *this; }
  QString& operator+=(QChar c) { d += c.toLatin1(); return *this; }
  QString& operator+=(const QByteArray& b) { d += b.d.c_str(); return *this; }
  QString& append(const QString& o) { d += o.d; return *this; }
  QString& append(QChar c) { d += c.toLatin1(); return *this; }
  QString& prepend(const QString& o) { d = o.d + d; return *this; }
  void push_back(const QString& o) { d += o.d; }
  void push_back(QChar c) { d += c.toLatin1(); }
  QByteArray toLatin1() const { return QByteArray(d); }
  QByteArray toUtf8() const { return QByteArray(d); }
  QByteArray toLocal8Bit() const { return QByteArray(d); }
#ifdef AERODDC_QSTRING_IS_STD_STRING
  // only for the build that puts the product's `vfo` class (std::string setters) under the reference's publisher.cpp
  operator std::string() const { return d; }
#endif
  std::string toStdString() const { return d; }
  static QString fromStdString(const std::string& s) { return QString(s); }
  static QString fromLatin1(const char* s, int n = -1) { return n < 0 ? QString(s) : QString(std::string(s, (size_t)n)); }
  static QString fromLatin1(const QByteArray& b) { return QString(b.d); }
  static QString fromUtf8(const char* s, int n = -1) { return fromLatin1(s, n); }
  static QString fromUtf8(const QByteArray& b) { return QString(b.d); }
  static QString fromLocal8Bit(const QByteArray& b) { return QString(b.d); }
  QString toUpper() const { std::string r = d; for (auto& c : r) c = (char)std::toupper((uchar)c); return QString(r); }
  QString toLower() const { std::string r = d; for (auto& c : r) c = (char)std::tolower((uchar)c); return QString(r); }
  QString trimmed() const { return QString(QByteArray(d).trimmed().d); }
  QString simplified() const {
    std::string r;
    bool sp = false;
    for (char c : QByteArray(d).trimmed().d) {
      if (std::isspace((uchar)c)) { sp = true; continue; }
      if (sp) r += ' ';
      sp = false;
      r += c;
    }
    return QString(r);
  }
  QString mid(int pos, int len = -1) const { return QString(QByteArray(d).mid(pos, len).d); }
  QString left(int n) const { return QString(QByteArray(d).left(n).d); }
  QString right(int n) const { return QString(QByteArray(d).right(n).d); }
  QString& remove(int pos, int n) { if (pos >= 0 && pos < size()) d.erase((size_t)pos, (size_t)n); return *this; }
  QString& remove(QChar c) { d.erase(std::remove(d.begin(), d.end(), c.toLatin1()), d.end()); return *this; }
  QString& remove(const QString& s) {
    if (s.d.empty()) return *this;
    size_t p;
    while ((p = d.find(s.d)) != std::string::npos) d.erase(p, s.d.size());
    return *this;
  }
  QString& replace(const QString& a, const QString& b) {
    if (a.d.empty()) return *this;
    size_t p = 0;
    while ((p = d.find(a.d, p)) != std::string::npos) { d.replace(p, a.d.size(), b.d); p += b.d.size(); }
    return *this;
  }
  QString& replace(QChar a, QChar b) { std::replace(d.begin(), d.end(), a.toLatin1(), b.toLatin1()); return *this; }
  QString& insert(int pos, const QString& s) { d.insert((size_t)pos, s.d); return *this; }
  bool contains(const QString& s) const { return d.find(s.d) != std::string::npos; }
  bool contains(QChar c) const { return d.find(c.toLatin1()) != std::string::npos; }
  bool startsWith(const QString& s) const { return d.compare(0, s.d.size(), s.d) == 0; }
  bool endsWith(const QString& s) const { return d.size() >= s.d.size() && d.compare(d.size() - s.d.size(), s.d.size(), s.d) == 0; }
  int indexOf(const QString& s, int from = 0) const { size_t p = d.find(s.d, (size_t)from); return p == std::string::npos ? -1 : (int)p; }
  int indexOf(QChar c, int from = 0) const { size_t p = d.find(c.toLatin1(), (size_t)from); return p == std::string::npos ? -1 : (int)p; }
  int lastIndexOf(QChar c) const { size_t p = d.rfind(c.toLatin1()); return p == std::string::npos ? -1 : (int)p; }
  int compare(const QString& o) const { return d.compare(o.d); }
  int toInt(bool* ok = nullptr, int base = 10) const { return QByteArray(d).toInt(ok, base); }
  uint toUInt(bool* ok = nullptr, int base = 10) const {
    char* e = nullptr;
    unsigned long v = std::strtoul(d.c_str(), &e, base);
    if (ok) *ok = e && *e == 0 && !d.empty();
    return (uint)v;
  }
  double toDouble(bool* ok = nullptr) const {
    char* e = nullptr;
    double v = std::strtod(d.c_str(), &e);
    if (ok) *ok = e && *e == 0 && !d.empty();
    return v;
  }
  QStringList split(const QString& sep) const;
  QStringList split(QChar sep) const;
  static QString vasprintf(const char* fmt, va_list ap) {
    va_list ap2;
    va_copy(ap2, ap);
    int n = std::vsnprintf(nullptr, 0, fmt, ap2);
    va_end(ap2);
    std::string r((size_t)(n < 0 ? 0 : n), 0);
    if (n > 0) std::vsnprintf(&r[0], (size_t)n + 1, fmt, ap);
    return QString(r);
  }
  static QString asprintf(const char* fmt, ...) __attribute__((format(printf, 1, 2))) {
    va_list ap;
    va_start(ap, fmt);
    QString r = vasprintf(fmt, ap);
    va_end(ap);
    return r;
  }
  QString& sprintf(const char* fmt, ...) __attribute__((format(printf, 2, 3))) {
    va_list ap;
    va_start(ap, fmt);
    *this = vasprintf(fmt, ap);
    va_end(ap);
    return *this;
  }
  static QString number(long long v, int base = 10) {
    if (base == 10) return QString(std::to_string(v));
    std::string r;
    bool neg = v < 0;
    unsigned long long u = neg ? (unsigned long long)(-v) : (unsigned long long)v;
    if (u == 0) r = "0";
    while (u) { r.insert(r.begin(), "0123456789abcdefghijklmnopqrstuvwxyz"[u % (unsigned)base]); u /= (unsigned)base; }
    return QString(neg ? "-" + r : r);
  }
  static QString number(int v, int base = 10) { return number((long long)v, base); }
  static QString number(uint v, int base = 10) { return number((long long)v, base); }
  static QString number(long v, int base = 10) { return number((long long)v, base); }
  static QString number(unsigned long v, int base = 10) { return number((long long)v, base); }
  static QString number(double v, char f = 'g', int prec = 6) {
    char fmt[16], b[400];
    std::snprintf(fmt, sizeof fmt, "%%.%d%c", prec, f);
    std::snprintf(b, sizeof b, fmt, v);
    return QString(b);
  }
  // %1..%99 place markers: the lowest-numbered marker is replaced (all its occurrences)
  QString argStr(const std::string& rep) const {
    int lowest = 100;
    for (size_t i = 0; i + 1 < d.size(); i++)
      if (d[i] == '%' && std::isdigit((uchar)d[i + 1])) {
        int n = d[i + 1] - '0';
        if (i + 2 < d.size() && std::isdigit((uchar)d[i + 2])) n = n * 10 + (d[i + 2] - '0');
        if (n > 0 && n < lowest) lowest = n;
      }
    if (lowest == 100) return *this;
    std::string r;
    for (size_t i = 0; i < d.size();) {
      if (d[i] == '%' && i + 1 < d.size() && std::isdigit((uchar)d[i + 1])) {
        int n = d[i + 1] - '0';
        size_t len = 2;
        if (i + 2 < d.size() && std::isdigit((uchar)d[i + 2])) { n = n * 10 + (d[i + 2] - '0'); len = 3; }
        if (n == lowest) { r += rep; i += len; continue; }
      }
      r += d[i++];
    }
    return QString(r);
  }
  static std::string pad(std::string s, int width, QChar fill) {
    int w = width < 0 ? -width : width;
    if ((int)s.size() < w) {
      std::string p((size_t)(w - (int)s.size()), fill.toLatin1());
      s = width < 0 ? s + p : p + s;
    }
    return s;
  }
  QString arg(const QString& a, int width = 0, QChar fill = QChar(' ')) const { return argStr(pad(a.d, width, fill)); }
  QString arg(const char* a, int width = 0, QChar fill = QChar(' ')) const { return argStr(pad(a, width, fill)); }
  QString arg(const QByteArray& a, int width = 0, QChar fill = QChar(' ')) const { return argStr(pad(a.d.c_str(), width, fill)); }
  QString arg(long long a, int width = 0, int base = 10, QChar fill = QChar(' ')) const { return argStr(pad(number(a, base).d, width, fill)); }
  QString arg(int a, int width = 0, int base = 10, QChar fill = QChar(' ')) const { return arg((long long)a, width, base, fill); }
  QString arg(uint a, int width = 0, int base = 10, QChar fill = QChar(' ')) const { return arg((long long)a, width, base, fill); }
  QString arg(long a, int width = 0, int base = 10, QChar fill = QChar(' ')) const { return arg((long long)a, width, base, fill); }
  QString arg(unsigned long a, int width = 0, int base = 10, QChar fill = QChar(' ')) const { return arg((long long)a, width, base, fill); }
  QString arg(uchar a, int width = 0, int base = 10, QChar fill = QChar(' ')) const { return arg((long long)a, width, base, fill); }
  QString arg(ushort a, int width = 0, int base = 10, QChar fill = QChar(' ')) const { return arg((long long)a, width, base, fill); }
  QString arg(short a, int width = 0, int base = 10, QChar fill = QChar(' ')) const { return arg((long long)a, width, base, fill); }
  QString arg(char a, int width = 0, QChar fill = QChar(' ')) const { return argStr(pad(std::string(1, a), width, fill)); }
  QString arg(QChar a, int width = 0, QChar fill = QChar(' ')) const { return argStr(pad(std::string(1, a.toLatin1()), width, fill)); }
  QString arg(double a, int width = 0, char f = 'g', int prec = -1, QChar fill = QChar(' ')) const {
    return argStr(pad(number(a, f, prec < 0 ? 6 : prec).d, width, fill));
  }
  bool operator==(const QString& o) const { return d == o.d; }
  bool operator!=(const QString& o) const { return d != o.d; }
  bool operator<(const QString& o) const { return d < o.d; }
  bool operator==(const char* s) const { return d == s; }
  bool operator!=(const char* s) const { return d != s; }
  std::string::const_iterator begin() const { return d.begin(); }
  std::string::const_iterator end() const { return d.end(); }
  std::string d;
};
inline QString operator+(const QString& a, const QString& b) { QString r(a); r += b; return r; }
inline QString operator+(const QString& a, const char* b) { QString r(a); r += b; return r; }
inline QString operator+(const char* a, const QString& b) { QString r(a); r += b; return r; }
inline QString operator+(const QString& a, char b) { QString r(a); r += b; return r; }
inline QString operator+(const QString& a, QChar b) { QString r(a); r += b; return r; }
inline QString operator+(const QString& a, const QByteArray& b) { QString r(a); r += b; return r; }
inline QByteArray& QByteArray::operator+=(const QString& s) { d += s.d; return *this; }
typedef QString QLatin1String;

class QStringList : public QList<QString> {
public:
  using QList<QString>::QList;
  QStringList() {}
  QString join(const QString& sep) const {
    QString r;
    for (int i = 0; i < size(); i++) { if (i) r += sep; r += (*this)[i]; }
    return r;
  }
  QStringList& operator<<(const QString& s) { push_back(s); return *this; }
};
inline QStringList QString::split(const QString& sep) const {
  QStringList r;
  if (sep.d.empty()) { r.push_back(*this); return r; }
  size_t p = 0, q;
  while ((q = d.find(sep.d, p)) != std::string::npos) { r.push_back(QString(d.substr(p, q - p))); p = q + sep.d.size(); }
  r.push_back(QString(d.substr(p)));
  return r;
}
inline QStringList QString::split(QChar sep) const { return split(QString(sep)); }

template <class K, class V> class QMap {
public:
  void clear() { m.clear(); }
  int count(const K& k) const { return (int)m.count(k); }
  int size() const { return (int)m.size(); }
  bool contains(const K& k) const { return m.count(k) != 0; }
  V value(const K& k, const V& def = V()) const { auto it = m.find(k); return it == m.end() ? def : it->second; }
  V take(const K& k) {
    auto it = m.find(k);
    if (it == m.end()) return V();
    V v = it->second;
    m.erase(it);
    return v;
  }
  int remove(const K& k) { return (int)m.erase(k); }
  void insert(const K& k, const V& v) { m[k] = v; }
  V& operator[](const K& k) { return m[k]; }
  QList<V> values() const { QList<V> r; for (auto& kv : m) r.push_back(kv.second); return r; }
  QList<K> keys() const { QList<K> r; for (auto& kv : m) r.push_back(kv.first); return r; }
  std::map<K, V> m;
};
template <class K, class V> class QCache {};

// qDebug() << ... : evaluated and dropped
class QDebug {
public:
  template <class T> QDebug& operator<<(const T&) { return *this; }
  QDebug& noquote() { return *this; }
  QDebug& nospace() { return *this; }
};
inline QDebug qDebug() { return QDebug(); }
inline QDebug qWarning() { return QDebug(); }
inline QDebug qCritical() { return QDebug(); }
inline QDebug qInfo() { return QDebug(); }

class QTimerEvent {
public:
  int timerId() const { return 0; }
};
class QEvent {};

namespace Qt {
enum ConnectionType { AutoConnection, DirectConnection, QueuedConnection, UniqueConnection = 0x80 };
enum DateFormat { TextDate, ISODate, ISODateWithMs };
enum TimeSpec { LocalTime, UTC };
}

class QObject {
public:
  explicit QObject(QObject* parent = nullptr) : parent_(parent) {}
  virtual ~QObject() {}
  QObject* parent() const { return parent_; }
  void setParent(QObject* p) { parent_ = p; }
  // every signal is routed by hand in the harness (it plays moc's part), so connections are accepted and ignored
  static bool connect(const QObject*, const char*, const QObject*, const char*, int = 0) { return true; }
  static bool disconnect(const QObject*, const char*, const QObject*, const char*) { return true; }
  bool disconnect() { return true; }
  int startTimer(int) { return 0; }
  void killTimer(int) {}
  void deleteLater() {}
  QObject* sender() const { return nullptr; }
  void setObjectName(const QString&) {}
protected:
  virtual void timerEvent(QTimerEvent*) {}
private:
  QObject* parent_;
};

template <class T> class QPointer {
public:
  QPointer() : p(nullptr) {}
  QPointer(T* q) : p(q) {}
  QPointer& operator=(T* q) { p = q; return *this; }
  T* operator->() const { return p; }
  T& operator*() const { return *p; }
  operator T*() const { return p; }
  bool isNull() const { return p == nullptr; }
  T* data() const { return p; }
  void clear() { p = nullptr; }
private:
  T* p;
};

class QTimer : public QObject {
public:
  explicit QTimer(QObject* parent = nullptr) : QObject(parent) {}
  void start(int = 0) {}
  void stop() {}
  void setInterval(int) {}
  void setSingleShot(bool) {}
  bool isActive() const { return false; }
};

class QElapsedTimer {
public:
  void start() {}
  qint64 restart() { return 0; }
  qint64 elapsed() const { return 0; }
  bool isValid() const { return true; }
  void invalidate() {}
};

class QThread : public QObject {
public:
  void start() {}
  void quit() {}
  bool wait(unsigned long = 0) { return true; }
};

class QDate {
public:
  QString toString(const QString&) const { return QString("00000000"); }
};
class QTime {
public:
  QString toString(const QString&) const { return QString("000000"); }
};
class QDateTime {
public:
  static QDateTime currentDateTime() { return QDateTime(); }
  static QDateTime currentDateTimeUtc() { return QDateTime(); }
  static qint64 currentMSecsSinceEpoch() { return 0; }
  static qint64 currentSecsSinceEpoch() { return 0; }
  qint64 toMSecsSinceEpoch() const { return 0; }
  qint64 toSecsSinceEpoch() const { return 0; }
  uint toTime_t() const { return 0; }
  QDateTime toUTC() const { return *this; }
  QDate date() const { return QDate(); }
  QTime time() const { return QTime(); }
  QString toString(const QString& = QString()) const { return QString("1970-01-01T00:00:00Z"); }
  QString toString(Qt::DateFormat) const { return QString("1970-01-01T00:00:00Z"); }
  qint64 secsTo(const QDateTime&) const { return 0; }
  qint64 msecsTo(const QDateTime&) const { return 0; }
  QDateTime addSecs(qint64) const { return *this; }
  bool operator<(const QDateTime&) const { return false; }
  bool operator>(const QDateTime&) const { return false; }
};

class QJsonObject {};
class QFile {};
class QTextStream {};

class QMetaEnum {
public:
  template <class T> static QMetaEnum fromType() { return QMetaEnum(); }
  // the only enum the sources ask about is DataBaseTextUser::DataBaseSchema (7 keys, databasetext.h)
  int keyCount() const { return 7; }
};

#endif
