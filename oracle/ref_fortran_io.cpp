// ref_fortran_io.cpp — TEST INFRASTRUCTURE (oracle/): a driver around the REFERENCE's own Fortran record IO class.
//
// The reference's Phantom reader / writer (shammodels/sph/src/io/PhantomDump.cpp) cannot be compiled here (it pulls
// SYCL vector types and the logging / unit libraries), but the record layer under it is header-only standard C++:
// shambase/include/shambase/fortran_io.hpp (FortranIOFile, load_fortran_file).  This driver is compiled against
// that header WHERE IT LIES under /root/reference (oracle/Makefile, target _ref/fortran_io_ref; nothing of the
// reference is copied) and walks a Phantom dump with it in the order PhantomDump::from_file / gen_file do
// (PhantomDump.cpp:276-375): every record of the file goes through the reference's typed read_* / write_* calls.
//
//   fortran_io_ref copy  IN OUT   read IN with FortranIOFile, write it again with FortranIOFile  (phantom_read_test.cpp)
//   fortran_io_ref write OUT      a synthetic dump with all eight element types and two blocks, written with
//                                 FortranIOFile (tests/test_io_formats.py builds the same content with
//                                 oracle/io_formats.py and compares the bytes)
//
// The two symbols below are the error-path hooks of shambase/exception.hpp, defined in shambase/src/exception.cpp
// (which needs the external fmt library): only reached when a record is malformed.
#include "shambase/fortran_io.hpp"
#include <array>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <string>
#include <vector>

namespace shambase {
    std::string exception_format(SourceLocation) { return ""; }
    void exception_gen_callback(std::string) {}
} // namespace shambase

using shambase::FortranIOFile;

template<class T>
static void copy_table(FortranIOFile &in, FortranIOFile &out) { // PhantomDumpTableHeader<T>::from_file / write
    int nvars;
    in.read(nvars);
    out.write(nvars);
    if (nvars == 0)
        return;
    std::vector<std::string> tags;
    in.read_string_array(tags, 16, nvars);
    std::vector<T> vals;
    in.read_val_array(vals, nvars);
    out.write_string_array(tags, 16, nvars);
    out.write_val_array(vals, nvars);
}
template<class T>
static void copy_arrays(FortranIOFile &in, FortranIOFile &out, i64 tot, int count) { // PhantomDumpBlockArray<T>
    for (int j = 0; j < count; j++) {
        std::string tag;
        in.read_fixed_string(tag, 16);
        std::vector<T> vals;
        in.read_val_array(vals, tot);
        out.write_fixed_string(tag, 16);
        out.write_val_array(vals, tot);
    }
}

static int do_copy(const std::string &fin, const std::string &fout) {
    FortranIOFile in = shambase::load_fortran_file(fin);
    FortranIOFile out;
    int i1, i2, iversion, i3;
    double r1;
    in.read(i1, r1, i2, iversion, i3);
    out.write(i1, r1, i2, iversion, i3);
    std::string fileid;
    in.read_fixed_string(fileid, 100);
    out.write_fixed_string(fileid, 100);
    copy_table<int>(in, out);
    copy_table<i8>(in, out);
    copy_table<i16>(in, out);
    copy_table<i32>(in, out);
    copy_table<i64>(in, out);
    copy_table<f64>(in, out);
    copy_table<f32>(in, out);
    copy_table<f64>(in, out);
    int nblocks;
    in.read(nblocks);
    out.write(nblocks);
    std::vector<i64> tots(nblocks);
    std::vector<std::array<i32, 8>> counts(nblocks);
    for (int b = 0; b < nblocks; b++) {
        in.read(tots[b], counts[b]);
        out.write(tots[b], counts[b]);
    }
    for (int b = 0; b < nblocks; b++) {
        copy_arrays<int>(in, out, tots[b], counts[b][0]);
        copy_arrays<i8>(in, out, tots[b], counts[b][1]);
        copy_arrays<i16>(in, out, tots[b], counts[b][2]);
        copy_arrays<i32>(in, out, tots[b], counts[b][3]);
        copy_arrays<i64>(in, out, tots[b], counts[b][4]);
        copy_arrays<f64>(in, out, tots[b], counts[b][5]);
        copy_arrays<f32>(in, out, tots[b], counts[b][6]);
        copy_arrays<f64>(in, out, tots[b], counts[b][7]);
    }
    if (!in.finished_read()) {
        std::cerr << "some data was not read\n";
        return 3;
    }
    out.write_to_file(fout);
    return 0;
}

// ---- the synthetic dump: table t has t + 1 entries "tag_<t>_<k>" = (k + 1) * (t + 2) (floats: x 0.5);
//      block 0: 37 values, {1,1,1,1,1,2,2,1} arrays; block 1: 3 values, three fort_real arrays;
//      array "arr_<b>_<t>_<j>"[i] = ((7 i + 3 j + t) mod 101) - 50 (floats: x 0.25)
static std::string pad(std::string s, size_t n) {
    s.resize(n, ' ');
    return s;
}
template<class T>
static void write_table(FortranIOFile &out, int t, bool is_float) {
    int nvars = t + 1;
    out.write(nvars);
    std::vector<std::string> tags;
    std::vector<T> vals;
    for (int k = 0; k < nvars; k++) {
        tags.push_back(pad("tag_" + std::to_string(t) + "_" + std::to_string(k), 16));
        double v = double((k + 1) * (t + 2)) * (is_float ? 0.5 : 1.0);
        vals.push_back(T(v));
    }
    out.write_string_array(tags, 16, nvars);
    out.write_val_array(vals, nvars);
}
template<class T>
static void write_arrays(FortranIOFile &out, int b, int t, i64 tot, int count, bool is_float) {
    for (int j = 0; j < count; j++) {
        std::string tag = pad("arr_" + std::to_string(b) + "_" + std::to_string(t) + "_" + std::to_string(j), 16);
        std::vector<T> vals;
        for (i64 i = 0; i < tot; i++) {
            double v = double((7 * i + 3 * j + t) % 101 - 50) * (is_float ? 0.25 : 1.0);
            vals.push_back(T(v));
        }
        out.write_fixed_string(tag, 16);
        out.write_val_array(vals, tot);
    }
}
static int do_write(const std::string &fout) {
    FortranIOFile out;
    int i1 = 60769, i2 = 60878, iversion = 1, i3 = 690706;
    double r1 = i2;
    out.write(i1, r1, i2, iversion, i3);
    std::string fileid = pad("FT:Phantom reference-IO pin", 100);
    out.write_fixed_string(fileid, 100);
    write_table<int>(out, 0, false);
    write_table<i8>(out, 1, false);
    write_table<i16>(out, 2, false);
    write_table<i32>(out, 3, false);
    write_table<i64>(out, 4, false);
    write_table<f64>(out, 5, true);
    write_table<f32>(out, 6, true);
    write_table<f64>(out, 7, true);
    int nblocks = 2;
    out.write(nblocks);
    i64 tots[2]                  = {37, 3};
    std::array<i32, 8> counts[2] = {{1, 1, 1, 1, 1, 2, 2, 1}, {0, 0, 0, 0, 0, 3, 0, 0}};
    for (int b = 0; b < 2; b++)
        out.write(tots[b], counts[b]);
    for (int b = 0; b < 2; b++) {
        write_arrays<int>(out, b, 0, tots[b], counts[b][0], false);
        write_arrays<i8>(out, b, 1, tots[b], counts[b][1], false);
        write_arrays<i16>(out, b, 2, tots[b], counts[b][2], false);
        write_arrays<i32>(out, b, 3, tots[b], counts[b][3], false);
        write_arrays<i64>(out, b, 4, tots[b], counts[b][4], false);
        write_arrays<f64>(out, b, 5, tots[b], counts[b][5], true);
        write_arrays<f32>(out, b, 6, tots[b], counts[b][6], true);
        write_arrays<f64>(out, b, 7, tots[b], counts[b][7], true);
    }
    out.write_to_file(fout);
    return 0;
}

int main(int argc, char **argv) {
    try {
        if (argc == 4 && std::string(argv[1]) == "copy")
            return do_copy(argv[2], argv[3]);
        if (argc == 3 && std::string(argv[1]) == "write")
            return do_write(argv[2]);
    } catch (const std::exception &e) {
        std::cerr << "fortran_io_ref: " << e.what() << "\n";
        return 2;
    }
    std::cerr << "usage: fortran_io_ref copy IN OUT | write OUT\n";
    return 1;
}
