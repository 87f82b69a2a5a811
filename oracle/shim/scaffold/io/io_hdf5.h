// TEST INFRASTRUCTURE ONLY -- stand-in for include/scaffold/io/io_hdf5.h (HDF5 C++ is absent
// from this image).  Same class name and member templates as the reference's IO
// (io_hdf5.h:9-212) so /root/reference/src compiles unchanged against it.
//  * Load()+Read(): pair-action tables come from a flat "PTAB1" container written by
//    simpimc_b200/tables.py (datasets keep the HDF5 names and C-order shapes of App. C of
//    SURVEY.md; bytes are copied in file order into the destination, which is what
//    H5::DataSet::read does for the reference).
//  * Create()/Write()/CreateExtendableDataSet()/AppendDataSet(): captured in memory so the
//    driver can hand block series (energies, g(r), S(k)) back to Python.
#ifndef ORACLE_SHIM_SCAFFOLD_IO_IO_HDF5_H_
#define ORACLE_SHIM_SCAFFOLD_IO_IO_HDF5_H_

#include <complex>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <string>
#include <vector>
#include <scaffold/matrix/matrix.h>

namespace scaffold { namespace io {

struct PtabDataset {
    int dtype;  // 0 f64, 1 u32, 2 i32, 3 string
    std::vector<uint64_t> dims;
    std::vector<char> bytes;
};

struct PtabFile {
    std::map<std::string, PtabDataset> sets;
    bool Load(const std::string &fn) {
        std::ifstream f(fn, std::ios::binary);
        if (!f) return false;
        char magic[6];
        f.read(magic, 6);
        if (std::memcmp(magic, "PTAB1\n", 6) != 0) return false;
        uint32_t n;
        f.read((char *)&n, 4);
        for (uint32_t i = 0; i < n; ++i) {
            uint32_t len;
            f.read((char *)&len, 4);
            std::string name(len, ' ');
            f.read(&name[0], len);
            PtabDataset ds;
            uint8_t dt;
            f.read((char *)&dt, 1);
            ds.dtype = dt;
            uint32_t nd;
            f.read((char *)&nd, 4);
            ds.dims.resize(nd);
            uint64_t count = 1;
            for (uint32_t k = 0; k < nd; ++k) {
                f.read((char *)&ds.dims[k], 8);
                count *= ds.dims[k];
            }
            uint64_t esz = (dt == 0) ? 8 : (dt == 3 ? 1 : 4);
            ds.bytes.resize(count * esz);
            f.read(ds.bytes.data(), ds.bytes.size());
            sets[name] = ds;
        }
        return true;
    }
};

inline std::string PtabKey(const std::string &name) {
    size_t i = 0;
    while (i < name.size() && name[i] == '/') ++i;
    std::string out;
    for (; i < name.size(); ++i) {
        if (name[i] == '/' && !out.empty() && out.back() == '/') continue;
        out.push_back(name[i]);
    }
    return out;
}

/// Everything the reference wrote, keyed by output file name then dataset name; each entry
/// is the list of appended records flattened to doubles.
inline std::map<std::string, std::map<std::string, std::vector<std::vector<double>>>> &CaptureStore() {
    static std::map<std::string, std::map<std::string, std::vector<std::vector<double>>>> store;
    return store;
}

class IO {
   public:
    std::string file_name;
    void Load(std::string &tmp_file_name) { file_name = tmp_file_name; }
    void Create() { CaptureStore()[file_name].clear(); }
    void CreateGroup(const std::string &) {}

    // ---- reading tables ----------------------------------------------------------------
    const PtabDataset &Find(const std::string &dataset_name) {
        static std::map<std::string, PtabFile> cache;
        auto it = cache.find(file_name);
        if (it == cache.end()) {
            PtabFile pf;
            if (!pf.Load(file_name)) {
                std::cerr << "oracle IO shim: cannot load table container " << file_name << std::endl;
                std::abort();
            }
            it = cache.insert(std::make_pair(file_name, pf)).first;
        }
        auto ds = it->second.sets.find(PtabKey(dataset_name));
        if (ds == it->second.sets.end()) {
            std::cerr << "oracle IO shim: dataset " << dataset_name << " not in " << file_name << std::endl;
            std::abort();
        }
        return ds->second;
    }
    static void CopyOut(const PtabDataset &ds, void *dst, size_t elem_size, size_t n_elem, int want_dtype) {
        if (ds.dtype != want_dtype || ds.bytes.size() != elem_size * n_elem) {
            std::cerr << "oracle IO shim: dataset type/size mismatch (" << ds.bytes.size() << " vs " << elem_size * n_elem << ")" << std::endl;
            std::abort();
        }
        std::memcpy(dst, ds.bytes.data(), ds.bytes.size());
    }
    void Read(const std::string &n, double &v) { CopyOut(Find(n), &v, 8, 1, 0); }
    void Read(const std::string &n, uint32_t &v) { CopyOut(Find(n), &v, 4, 1, 1); }
    void Read(const std::string &n, int &v) { CopyOut(Find(n), &v, 4, 1, 2); }
    void Read(const std::string &n, matrix::vec<double> &v) { CopyOut(Find(n), v.memptr(), 8, v.size(), 0); }
    void Read(const std::string &n, matrix::mat<double> &v) { CopyOut(Find(n), v.memptr(), 8, v.size(), 0); }
    void Read(const std::string &n, matrix::cube<double> &v) { CopyOut(Find(n), v.memptr(), 8, v.size(), 0); }
    void Read(const std::string &n, std::string &v) {
        const PtabDataset &ds = Find(n);
        v.assign(ds.bytes.begin(), ds.bytes.end());
    }

    // ---- capturing output --------------------------------------------------------------
    static std::vector<double> Flat(const double &v) { return {v}; }
    static std::vector<double> Flat(const int &v) { return {(double)v}; }
    static std::vector<double> Flat(const unsigned int &v) { return {(double)v}; }
    static std::vector<double> Flat(const bool &v) { return {(double)v}; }
    static std::vector<double> Flat(const std::string &) { return {}; }
    template <class E>
    static std::vector<double> Flat(const matrix::vec<E> &v) {
        std::vector<double> o(v.size());
        for (size_t i = 0; i < v.size(); ++i) o[i] = (double)v(i);
        return o;
    }
    template <class E>
    static std::vector<double> Flat(const matrix::mat<E> &v) {
        std::vector<double> o(v.size());
        for (size_t i = 0; i < v.size(); ++i) o[i] = (double)v.memptr()[i];
        return o;
    }
    template <class E>
    static std::vector<double> Flat(const matrix::cube<E> &v) {
        std::vector<double> o(v.d.size());
        for (size_t i = 0; i < v.d.size(); ++i) o[i] = (double)v.d[i];
        return o;
    }
    template <class T>
    void Write(const std::string &dataset_name, T &data) {
        auto &rec = CaptureStore()[file_name][PtabKey(dataset_name)];
        rec.clear();
        rec.push_back(Flat(data));
    }
    template <class T>
    void Rewrite(const std::string &dataset_name, T &data) { Write(dataset_name, data); }
    template <class T>
    void CreateExtendableDataSet(const std::string &prefix, const std::string &dataset_name, T &data) {
        auto &rec = CaptureStore()[file_name][PtabKey(prefix + dataset_name)];
        rec.clear();
        rec.push_back(Flat(data));
    }
    template <class T>
    void AppendDataSet(const std::string &prefix, const std::string &dataset_name, T &data) {
        CaptureStore()[file_name][PtabKey(prefix + dataset_name)].push_back(Flat(data));
    }
};

}}  // namespace scaffold::io

#endif  // ORACLE_SHIM_SCAFFOLD_IO_IO_HDF5_H_
