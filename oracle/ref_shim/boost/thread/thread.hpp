/* boost/thread/thread.hpp — stand-in for Boost.Thread (not installed).  TEST INFRASTRUCTURE ONLY.
 * TreadedRenderer (cell/ppu_renderer.cpp:132-142) needs thread_group::create_thread(functor) and join_all(). */
#ifndef YV_REF_SHIM_BOOST_THREAD_HPP
#define YV_REF_SHIM_BOOST_THREAD_HPP
#include <thread>
#include <vector>
namespace boost {
class thread_group {
  std::vector<std::thread> m_threads;
public:
  template <class F> void create_thread(F f) { m_threads.emplace_back(f); }
  void join_all() { for (auto &t : m_threads) t.join(); m_threads.clear(); }
  ~thread_group() { join_all(); }
};
}
#endif
