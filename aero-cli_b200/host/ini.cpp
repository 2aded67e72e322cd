#include "ini.h"

#include <cstdlib>
#include <fstream>
#include <sstream>

namespace aero {

static std::string trim(const std::string& s) {
  size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
  return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
}

bool IniSettings::load(const std::string& path) {
  std::ifstream f(path);
  if (!f) return false;
  std::stringstream ss;
  ss << f.rdbuf();
  return loadString(ss.str());
}

bool IniSettings::loadString(const std::string& text) {
  kv_.clear();
  std::istringstream in(text);
  std::string line, group;
  while (std::getline(in, line)) {
    line = trim(line);
    if (line.empty() || line[0] == ';' || line[0] == '#') continue;
    if (line[0] == '[') {
      const size_t e = line.find(']');
      group = trim(line.substr(1, e == std::string::npos ? std::string::npos : e - 1));
      if (group == "General") group.clear();   // QSettings maps top-level keys to [General]
      continue;
    }
    const size_t eq = line.find('=');
    if (eq == std::string::npos) continue;
    std::string key = trim(line.substr(0, eq)), val = trim(line.substr(eq + 1));
    if (val.size() >= 2 && val.front() == '"' && val.back() == '"') val = val.substr(1, val.size() - 2);
    for (char& c : key)
      if (c == '\\') c = '/';                   // "1\frequency" is the key path 1/frequency
    kv_[group.empty() ? key : group + "/" + key] = val;
  }
  return true;
}

std::string IniSettings::scoped(const std::string& key) const {
  if (array_.empty()) return key;
  if (index_ < 0) return array_ + "/" + key;
  return array_ + "/" + std::to_string(index_ + 1) + "/" + key;
}

std::string IniSettings::value(const std::string& key) const {
  auto it = kv_.find(scoped(key));
  return it == kv_.end() ? std::string() : it->second;
}
bool IniSettings::contains(const std::string& key) const { return kv_.count(scoped(key)) != 0; }

int IniSettings::toInt(const std::string& key) const {
  const std::string v = value(key);
  if (v.empty()) return 0;
  char* end = nullptr;
  const long long i = std::strtoll(v.c_str(), &end, 10);
  if (end && *end == 0) return (int)i;
  // QVariant(QString).toInt() fails on "12.5" -> 0; a double-looking integer such as "1e3" also fails
  return 0;
}
float IniSettings::toFloat(const std::string& key) const {
  const std::string v = value(key);
  if (v.empty()) return 0.0f;
  char* end = nullptr;
  const float f = std::strtof(v.c_str(), &end);
  return (end && *end == 0) ? f : 0.0f;
}

int IniSettings::beginReadArray(const std::string& name) {
  array_ = name;
  index_ = -1;
  return toInt("size");
}
void IniSettings::setArrayIndex(int i) { index_ = i; }
void IniSettings::endArray() {
  array_.clear();
  index_ = -1;
}

}  // namespace aero
