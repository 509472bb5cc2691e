#include "config.hpp"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <fstream>

Config::Config(std::stringstream& logstream, bool quietmode_) : log(&logstream), quietmode(quietmode_) {}

void Config::echo(const std::string& key, const std::string& value, bool) const { (*log) << key << " = " << value << std::endl; }

void Config::ReadFile(const std::string& filename) {
  std::ifstream f(filename);
  if (!f.is_open()) {
    printf("ERROR: Unable to read the file %s\n", filename.c_str());
    exit(1);
  }
  std::string line;
  while (std::getline(f, line)) {
    line.erase(std::remove_if(line.begin(), line.end(), [](unsigned char c) { return std::isspace(c); }), line.end());
    if (line.empty() || line[0] == '#' || line[0] == '/') continue;
    size_t eq = line.find('=');
    if (eq == std::string::npos) continue;
    std::string key = line.substr(0, eq), value = line.substr(eq + 1);
    if (kv.count(key) && !quietmode) printf("# Warning: existing param found : %s, with new value %s.\n", key.c_str(), value.c_str());
    kv[key] = value;
  }
}

std::string Config::GetStrParam(const std::string& key, const std::string& dflt, bool, bool exportme) const {
  auto it = kv.find(key);
  std::string v = it == kv.end() ? dflt : it->second;
  if (exportme) echo(key, v, it == kv.end());
  return v;
}
int Config::GetIntParam(const std::string& key, int dflt, bool, bool exportme) const {
  auto it = kv.find(key);
  int v = it == kv.end() ? dflt : atoi(it->second.c_str());
  if (exportme) echo(key, std::to_string(v), it == kv.end());
  return v;
}
double Config::GetDoubleParam(const std::string& key, double dflt, bool, bool exportme) const {
  auto it = kv.find(key);
  double v = it == kv.end() ? dflt : atof(it->second.c_str());
  if (exportme) echo(key, std::to_string(v), it == kv.end());
  return v;
}
bool Config::GetBoolParam(const std::string& key, bool dflt, bool, bool exportme) const {
  auto it = kv.find(key);
  bool v = dflt;
  if (it != kv.end()) {
    std::string s = it->second;
    std::transform(s.begin(), s.end(), s.begin(), ::tolower);
    v = (s == "true" || s == "1" || s == "yes");
  }
  if (exportme) echo(key, v ? "true" : "false", it == kv.end());
  return v;
}
static std::vector<std::string> split_commas(const std::string& s) {
  std::vector<std::string> out;
  std::string item;
  std::stringstream ss(s);
  while (std::getline(ss, item, ',')) out.push_back(item);
  return out;
}
void Config::GetVecStrParam(const std::string& key, std::vector<std::string>& fillme, const std::string& dflt, bool, bool exportme) const {
  auto it = kv.find(key);
  fillme = split_commas(it == kv.end() ? dflt : it->second);
  if (fillme.empty()) fillme.push_back(dflt);
  if (exportme) echo(key, it == kv.end() ? dflt : it->second, it == kv.end());
}
void Config::GetVecDoubleParam(const std::string& key, std::vector<double>& fillme, double dflt, bool, bool exportme) const {
  fillme.clear();
  auto it = kv.find(key);
  if (it == kv.end()) fillme.push_back(dflt);
  else for (auto& s : split_commas(it->second)) fillme.push_back(atof(s.c_str()));
  if (fillme.empty()) fillme.push_back(dflt);
  if (exportme) echo(key, it == kv.end() ? std::to_string(dflt) : it->second, it == kv.end());
}
void Config::GetVecIntParam(const std::string& key, std::vector<int>& fillme, int dflt, bool, bool exportme) const {
  fillme.clear();
  auto it = kv.find(key);
  if (it == kv.end()) fillme.push_back(dflt);
  else for (auto& s : split_commas(it->second)) fillme.push_back(atoi(s.c_str()));
  if (fillme.empty()) fillme.push_back(dflt);
  if (exportme) echo(key, it == kv.end() ? std::to_string(dflt) : it->second, it == kv.end());
}
