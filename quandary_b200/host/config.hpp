/* config.hpp -- the reference's `key = value` configuration file (src/config.cpp:38-237, config_template.cfg): lines starting
 * with '#' or '/' and blank lines are ignored, all whitespace is stripped, duplicates overwrite, list values are comma-separated. */
#pragma once
#include <map>
#include <sstream>
#include <string>
#include <vector>

class Config {
  std::map<std::string, std::string> kv;
  std::stringstream* log;
  bool quietmode;
  void echo(const std::string& key, const std::string& value, bool warned_default) const;

 public:
  Config(std::stringstream& logstream, bool quietmode_ = false);
  void ReadFile(const std::string& filename);
  size_t count(const std::string& key) const { return kv.count(key); }
  void SetParam(const std::string& key, const std::string& value) { kv[key] = value; }

  std::string GetStrParam(const std::string& key, const std::string& dflt = "", bool exitonerror = true, bool warnme = true) const;
  int GetIntParam(const std::string& key, int dflt = -1, bool exitonerror = true, bool warnme = true) const;
  double GetDoubleParam(const std::string& key, double dflt = -1.0, bool exitonerror = true, bool warnme = true) const;
  bool GetBoolParam(const std::string& key, bool dflt = false, bool exitonerror = true, bool warnme = true) const;
  void GetVecStrParam(const std::string& key, std::vector<std::string>& fillme, const std::string& dflt = "none", bool exitonerror = true, bool warnme = true) const;
  void GetVecDoubleParam(const std::string& key, std::vector<double>& fillme, double dflt = 1e20, bool exitonerror = true, bool warnme = true) const;
  void GetVecIntParam(const std::string& key, std::vector<int>& fillme, int dflt = -1, bool exitonerror = true, bool warnme = true) const;
};
