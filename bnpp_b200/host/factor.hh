// bn::Factor -- dense fp64 potential table.  Public API of reference code/factor.hh:10-48.
//
// What changed underneath: the table lives in HBM.  Every operation is a thin wrapper
// over an extern "C" entry point of include/bnpp_b200.h and returns a new device-resident
// Factor without synchronising; the partition of a result stays on the device until
// partition() is asked for.  operator[] and operator<< read through a host mirror that is
// downloaded lazily (and uploaded again if written through the non-const operator[]).
// Ownership is the reference's: a Factor owns and deletes its Domain (code/factor.cpp:49-52).
#ifndef BNPP_HOST_FACTOR_HH
#define BNPP_HOST_FACTOR_HH

#include "domain.hh"

#include <vector>

namespace bn {

class Factor {
public:
    Factor(const Domain *domain, std::vector<double> values, double partition);
    Factor(const Domain *domain, double value = 0.0);
    Factor(double value = 1.0);
    Factor(const Factor &f);
    Factor(Factor &&f);
    ~Factor();

    Factor &operator=(Factor &&f);
    Factor operator*(const Factor &f);
    void operator*=(const Factor &f);

    const Domain &domain() const { return *_domain; }
    unsigned size()        const { return _domain->size();  }
    unsigned width()       const { return _domain->width(); }
    double partition()     const;

    const double &operator[](unsigned i) const;
    double &operator[](unsigned i);

    double max() const;
    double min() const;

    Factor sum_out(const Variable *variable) const;
    Factor product(const Factor &f) const;
    Factor divide(const Factor &f) const;
    Factor conditioning(const std::unordered_map<unsigned,unsigned> &evidence) const;
    Factor normalize() const;

    std::unordered_map<unsigned,unsigned> sampling(const std::unordered_map<unsigned,unsigned> &evidence) const;

    friend std::ostream &operator<<(std::ostream &os, const Factor &f);

    // ---- additions for the device path (not in the reference) ----
    const double *device_data() const;      // table in HBM (uploads a host-built factor on first use)
    bool resident() const { return _dev_valid; }
    // adopts a device buffer of size()+1 doubles (partition in the last slot) produced by the library
    static Factor adopt(const Domain *domain, double *dev);

private:
    struct Uninit {};
    Factor(const Domain *domain, Uninit);   // device buffer allocated, contents to be written by a kernel
    void release();
    void sync_host() const;
    double *dev_z() const { return _dev + size(); }

    const Domain *_domain;
    mutable double *_dev;                   // size()+1 doubles, or nullptr
    mutable std::vector<double> _host;
    mutable bool _dev_valid, _host_valid;
    mutable double _partition;
    mutable bool _z_pending;                // the partition is still only in _dev[size()]
};

}  // namespace bn

#endif
