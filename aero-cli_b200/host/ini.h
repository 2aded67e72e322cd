// Minimal reader for the QSettings IniFormat subset that aero-publish's settings files use
// (/root/reference/publish/publisher.cpp:55-227): top-level keys (QSettings' "General" group),
// and arrays written as
//     [vfos]
//     size=2
//     1\frequency=1545000000
//     1\topic=VFO01
// read with beginReadArray(name) / setArrayIndex(i) / value(key) / endArray().
#pragma once
#include <map>
#include <string>

namespace aero {

class IniSettings {
 public:
  // false if the file cannot be opened
  bool load(const std::string& path);
  bool loadString(const std::string& text);

  // value of `key` in the current scope ("" when absent, like an invalid QVariant's toString())
  std::string value(const std::string& key) const;
  int toInt(const std::string& key) const;       // QVariant::toInt: 0 when absent or not a number
  float toFloat(const std::string& key) const;   // QVariant::toFloat
  bool contains(const std::string& key) const;

  int beginReadArray(const std::string& name);   // returns the array's size entry
  void setArrayIndex(int i);                     // 0-based, as in Qt (stored 1-based in the file)
  void endArray();

 private:
  std::string scoped(const std::string& key) const;
  std::map<std::string, std::string> kv_;        // "group/key" or "General-less key"
  std::string array_;
  int index_ = -1;
};

}  // namespace aero
