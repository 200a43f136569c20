/* utils.h — stand-in for the reference's cpp/utils.h (absent).  TEST INFRASTRUCTURE ONLY.
 * read(stream, value): raw binary read, as cell/svodata.h:35-40 uses it. */
#ifndef YV_REF_SHIM_UTILS_H
#define YV_REF_SHIM_UTILS_H
template <class T>
inline void read(std::istream &in, T &v) { in.read(reinterpret_cast<char *>(&v), sizeof(T)); }
#endif
