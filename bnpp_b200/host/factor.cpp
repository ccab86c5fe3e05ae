#include "factor.hh"
#include "runtime.hh"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <iomanip>
#include <iostream>
#include <random>

namespace bn {

namespace {

bnpp_scope scope_of(const Domain &d)
{
    bnpp_scope s;
    s.rank = (int32_t)d.width();
    s.var_id = d.ids().data();
    s.card = d.cards().data();
    return s;
}

double *device_alloc(uint64_t n)
{
    double *p = nullptr;
    gpu::check(bnpp_alloc(gpu::ctx(), n, &p), "bnpp_alloc");
    return p;
}

}  // namespace

Factor::Factor(const Domain *domain, std::vector<double> values, double partition)
    : _domain(domain), _dev(nullptr), _host(std::move(values)), _dev_valid(false), _host_valid(true),
      _partition(partition), _z_pending(false)
{
}

Factor::Factor(const Domain *domain, double value)
    : _domain(domain), _dev(nullptr), _host(domain->size(), value), _dev_valid(false), _host_valid(true),
      _partition(domain->size() * value), _z_pending(false)
{
}

Factor::Factor(double value)
    : _domain(new Domain()), _dev(nullptr), _host(1, value), _dev_valid(false), _host_valid(true),
      _partition(value), _z_pending(false)
{
}

Factor::Factor(const Domain *domain, Uninit)
    : _domain(domain), _dev(device_alloc((uint64_t)domain->size() + 1)), _dev_valid(true), _host_valid(false),
      _partition(0.0), _z_pending(true)
{
}

Factor Factor::adopt(const Domain *domain, double *dev)
{
    Factor f(domain, std::vector<double>(), 0.0);
    f._dev = dev;
    f._dev_valid = true;
    f._host_valid = false;
    f._z_pending = true;
    return f;
}

// deep copy, Domain included (code/factor.cpp:32-37)
Factor::Factor(const Factor &f)
    : _domain(new Domain(*f._domain)), _dev(nullptr), _host(), _dev_valid(false), _host_valid(f._host_valid),
      _partition(f._partition), _z_pending(f._z_pending)
{
    if (f._host_valid) _host = f._host;
    if (f._dev_valid) {
        const uint64_t n = (uint64_t)size() + 1;
        _dev = device_alloc(n);
        // a contraction of one operand onto its own scope is a device-to-device copy
        bnpp_scope s = scope_of(*_domain);
        bnpp_operand op = {f._dev, s, nullptr};
        gpu::check(bnpp_product_sum_out(gpu::ctx(), 1, &op, &s, -1, 0, _dev, nullptr), "copy");
        // the partition is copied, not re-summed (the reference copies `_partition` verbatim)
        bnpp_scope scalar = {0, nullptr, nullptr};
        bnpp_operand z = {f.dev_z(), scalar, nullptr};
        if (f._z_pending) gpu::check(bnpp_product_sum_out(gpu::ctx(), 1, &z, &scalar, -1, 0, dev_z(), nullptr), "copy");
        else gpu::check(bnpp_fill(gpu::ctx(), dev_z(), 1, f._partition), "bnpp_fill");
        _dev_valid = true;
    }
}

Factor::Factor(Factor &&f)
    : _domain(f._domain), _dev(f._dev), _host(std::move(f._host)), _dev_valid(f._dev_valid), _host_valid(f._host_valid),
      _partition(f._partition), _z_pending(f._z_pending)
{
    f._domain = nullptr;
    f._dev = nullptr;
    f._dev_valid = f._host_valid = f._z_pending = false;
    f._partition = 0.0;
}

void Factor::release()
{
    if (_dev) bnpp_free(gpu::ctx(), _dev);
    _dev = nullptr;
    delete _domain;
    _domain = nullptr;
}

Factor::~Factor() { release(); }

Factor &Factor::operator=(Factor &&f)
{
    if (this != &f) {
        release();
        _domain = f._domain;
        _dev = f._dev;
        _host = std::move(f._host);
        _dev_valid = f._dev_valid;
        _host_valid = f._host_valid;
        _partition = f._partition;
        _z_pending = f._z_pending;
        f._domain = nullptr;
        f._dev = nullptr;
        f._dev_valid = f._host_valid = f._z_pending = false;
        f._partition = 0.0;
    }
    return *this;
}

Factor Factor::operator*(const Factor &f) { return product(f); }

void Factor::operator*=(const Factor &f) { *this = product(f); }

const double *Factor::device_data() const
{
    if (!_dev_valid) {
        const uint64_t n = size();
        // a table with fewer values than its scope implies (a malformed UAI file): the reference dies in
        // _values.at(); uploading n doubles from a shorter vector would read past the heap buffer
        if (_host.size() < n) throw "Factor: table holds fewer values than its domain.";
        if (!_dev) _dev = device_alloc(n + 1);
        gpu::check(bnpp_upload(gpu::ctx(), _dev, _host.data(), n), "bnpp_upload");
        gpu::check(bnpp_fill(gpu::ctx(), dev_z(), 1, _partition), "bnpp_fill");
        // the host vector may be modified or freed before the copy runs on the stream
        gpu::check(bnpp_ctx_sync(gpu::ctx()), "bnpp_ctx_sync");
        _dev_valid = true;
    }
    return _dev;
}

void Factor::sync_host() const
{
    if (_host_valid) return;
    _host.resize(size());
    gpu::check(bnpp_download(gpu::ctx(), _host.data(), _dev, size()), "bnpp_download");
    _host_valid = true;
}

double Factor::partition() const
{
    if (_z_pending) {
        gpu::check(bnpp_download(gpu::ctx(), &_partition, dev_z(), 1), "bnpp_download");
        _z_pending = false;
    }
    return _partition;
}

const double &Factor::operator[](unsigned i) const
{
    if (i >= size()) throw "Factor::operator[]: Index out of range.";   // code/factor.cpp:87
    sync_host();
    return _host[i];
}

// Writes go to the host mirror; the device copy is refreshed before its next use.  As in
// the reference the cached partition is NOT updated (code/factor.cpp:90-95).
double &Factor::operator[](unsigned i)
{
    if (i >= size()) throw "Factor::operator[]: Index out of range.";
    sync_host();
    partition();
    _dev_valid = false;
    return _host[i];
}

double Factor::max() const
{
    double r = 0.0;
    double *slot = device_alloc(1);
    gpu::check(bnpp_reduce(gpu::ctx(), 1, size(), device_data(), 0.0, slot), "bnpp_reduce");
    gpu::check(bnpp_download(gpu::ctx(), &r, slot, 1), "bnpp_download");
    bnpp_free(gpu::ctx(), slot);
    return r;
}

double Factor::min() const
{
    double r = 0.0;
    double *slot = device_alloc(1);
    gpu::check(bnpp_reduce(gpu::ctx(), 2, size(), device_data(), partition(), slot), "bnpp_reduce");
    gpu::check(bnpp_download(gpu::ctx(), &r, slot, 1), "bnpp_download");
    bnpp_free(gpu::ctx(), slot);
    return r;
}

Factor Factor::product(const Factor &f) const
{
    Factor out(new Domain(*_domain, *f._domain), Uninit());
    bnpp_scope sa = scope_of(*_domain), sb = scope_of(*f._domain);
    gpu::check(bnpp_product(gpu::ctx(), &sa, device_data(), &sb, f.device_data(), 0, out._dev, out.dev_z()), "bnpp_product");
    return out;
}

Factor Factor::divide(const Factor &f) const
{
    Factor out(new Domain(*_domain, *f._domain), Uninit());
    bnpp_scope sa = scope_of(*_domain), sb = scope_of(*f._domain);
    gpu::check(bnpp_product(gpu::ctx(), &sa, device_data(), &sb, f.device_data(), 1, out._dev, out.dev_z()), "bnpp_product");
    uint32_t bits = 0;
    gpu::check(bnpp_ctx_status(gpu::ctx(), &bits, 1), "bnpp_ctx_status");
    if (bits & BNPP_STATUS_ZERO_DIVISOR) {
        // the reference asserts per entry (code/factor.cpp:169, asserts are live in its release build)
        std::fprintf(stderr, "Factor::divide: Assertion `f[pos2] != 0' failed.\n");
        std::abort();
    }
    return out;
}

Factor Factor::sum_out(const Variable *variable) const
{
    if (!_domain->in_scope(variable)) return Factor(*this);   // code/factor.cpp:185-188
    Factor out(new Domain(*_domain, variable), Uninit());
    bnpp_scope s = scope_of(*_domain);
    gpu::check(bnpp_sum_out(gpu::ctx(), &s, device_data(), variable->id(), out._dev, out.dev_z()), "bnpp_sum_out");
    return out;
}

Factor Factor::conditioning(const std::unordered_map<unsigned,unsigned> &evidence) const
{
    Factor out(new Domain(*_domain, evidence), Uninit());
    std::vector<uint32_t> var, val;
    for (auto &e : evidence) {
        var.push_back(e.first);
        val.push_back(e.second);
    }
    bnpp_scope s = scope_of(*_domain);
    gpu::check(bnpp_condition(gpu::ctx(), &s, device_data(), (int)var.size(), var.data(), val.data(), out._dev, out.dev_z()),
               "bnpp_condition");
    return out;
}

// divides by the CACHED partition (code/factor.cpp:250); when that partition has not left
// the device yet the kernel reads it there, so a chain product -> normalize never syncs
Factor Factor::normalize() const
{
    Factor out(new Domain(*_domain), Uninit());
    const double *in = device_data();
    gpu::check(bnpp_normalize(gpu::ctx(), size(), in, _z_pending ? dev_z() : nullptr, _partition, out._dev), "bnpp_normalize");
    out._partition = 1.0;
    out._z_pending = false;
    return out;
}

// Out of the hot path (stochastic inference).  Same procedure as
// code/factor.cpp:257-288: condition, normalise if needed, inverse-CDF draw.
std::unordered_map<unsigned,unsigned> Factor::sampling(const std::unordered_map<unsigned,unsigned> &evidence) const
{
    std::random_device rd;
    double prob = rd();
    prob /= rd.max();
    Factor f = conditioning(evidence);
    if (std::fabs(f.partition() - 1.0) > 0.001) f = f.normalize();
    const Domain &d = f.domain();
    std::vector<unsigned> valuation(d.width(), 0);
    double p = 0.0;
    for (unsigned i = 0; i < f.size(); ++i) {
        p += f[i];
        if (prob <= p) break;
        d.next_valuation(valuation);
    }
    std::unordered_map<unsigned,unsigned> sample;
    for (unsigned i = 0; i < d.width(); ++i) sample[d[i]->id()] = valuation[i];
    return sample;
}

// Byte-compatible with code/factor.cpp:291-321, including the sticky `fixed <<
// setprecision(7)` it leaves on the stream (SURVEY §8b "CLI text").
std::ostream &operator<<(std::ostream &os, const Factor &f)
{
    const int width = f.width();
    const int size = f.size();
    os << "Factor(" << "width:" << width << ", " << "size:" << size << ", " << "partition:" << f.partition() << ")" << std::endl;
    for (int i = 0; i < width; ++i) os << f.domain()[i]->id() << " ";
    os << std::endl;
    std::vector<unsigned> valuation(width, 0);
    for (int i = 0; i < size; ++i) {
        for (int j = 0; j < width; ++j) os << valuation[j] << " ";
        os << ": " << std::fixed << std::setprecision(7) << f[i] << std::endl;
        f.domain().next_valuation(valuation);
    }
    return os;
}

}  // namespace bn
