/* nmpc_b200 -- functor registry storage. */
#include <nmpc_b200/engine/registry.h>

#include <map>
#include <mutex>

namespace nmpc_b200
{
namespace
{
struct Registry
{
  std::mutex mutex;
  std::map<std::string, std::unique_ptr<ModelEntry>> entries;
  std::vector<std::string> names;
};

Registry & registry()
{
  static Registry * r = new Registry(); // never destroyed: registrars of other TUs may outlive statics
  return *r;
}
} // namespace

ModelEntry & registryEntry(const std::string & name)
{
  Registry & r = registry();
  std::lock_guard<std::mutex> lock(r.mutex);
  auto it = r.entries.find(name);
  if(it == r.entries.end())
  {
    auto e = std::make_unique<ModelEntry>();
    e->name = name;
    it = r.entries.emplace(name, std::move(e)).first;
    r.names.push_back(name);
  }
  return *it->second;
}

const ModelEntry * registryFind(const std::string & name)
{
  Registry & r = registry();
  std::lock_guard<std::mutex> lock(r.mutex);
  auto it = r.entries.find(name);
  return it == r.entries.end() ? nullptr : it->second.get();
}

const std::vector<std::string> & registryNames()
{
  return registry().names;
}
} // namespace nmpc_b200
