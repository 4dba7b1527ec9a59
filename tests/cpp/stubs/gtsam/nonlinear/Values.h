#pragma once
#include "../base.h"
namespace gtsam {
// at<T>(key) throws when the key is missing, like gtsam::Values (the reference relies on that for G(0),
// geometric_factor.hpp:257)
class Values {
 public:
  template <typename T>
  void insert(Key k, const T& v) {
    slots_[k] = Slot{std::type_index(typeid(T)), std::make_shared<T>(v)};
  }
  template <typename T>
  const T& at(Key k) const {
    auto it = slots_.find(k);
    if (it == slots_.end()) throw std::out_of_range("ValuesKeyDoesNotExist");
    if (it->second.type != std::type_index(typeid(T))) throw std::invalid_argument("ValuesIncorrectType");
    return *static_cast<const T*>(it->second.p.get());
  }

 private:
  struct Slot {
    std::type_index type = std::type_index(typeid(void));
    std::shared_ptr<void> p;
  };
  std::map<Key, Slot> slots_;
};
}  // namespace gtsam
