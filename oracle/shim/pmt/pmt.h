/* TEST INFRASTRUCTURE ONLY (oracle shim) - stands in for <pmt/pmt.h>: the polymorphic-type subset the
 * reference's blocks use (symbols for port names, long integers for events, std::any payloads). */
#pragma once
#include <any>
#include <memory>
#include <stdexcept>
#include <string>
namespace pmt
{
struct pmt_base
{
    enum kind_t
    {
        K_NULL,
        K_SYMBOL,
        K_LONG,
        K_ANY
    } kind{K_NULL};
    std::string sym;
    long lval{0};
    std::any aval;
};
typedef std::shared_ptr<pmt_base> pmt_t;

inline pmt_t intern(const std::string& s)
{
    auto p = std::make_shared<pmt_base>();
    p->kind = pmt_base::K_SYMBOL;
    p->sym = s;
    return p;
}
inline pmt_t string_to_symbol(const std::string& s) { return intern(s); }
inline std::string symbol_to_string(const pmt_t& p)
{
    if (!p || p->kind != pmt_base::K_SYMBOL) throw std::runtime_error("pmt: not a symbol");
    return p->sym;
}
inline bool is_symbol(const pmt_t& p) { return p && p->kind == pmt_base::K_SYMBOL; }
inline pmt_t from_long(long v)
{
    auto p = std::make_shared<pmt_base>();
    p->kind = pmt_base::K_LONG;
    p->lval = v;
    return p;
}
inline bool is_integer(const pmt_t& p) { return p && p->kind == pmt_base::K_LONG; }
inline long to_long(const pmt_t& p)
{
    if (!is_integer(p)) throw std::runtime_error("pmt: not an integer");
    return p->lval;
}
inline pmt_t make_any(const std::any& a)
{
    auto p = std::make_shared<pmt_base>();
    p->kind = pmt_base::K_ANY;
    p->aval = a;
    return p;
}
inline bool is_any(const pmt_t& p) { return p && p->kind == pmt_base::K_ANY; }
/* GNU Radio stores integers published with from_long as pmt integers; blocks that read a message with
 * any_ref expect an any.  The reference publishes its events with pmt::make_any(int) where it reads them
 * with any_ref, and with from_long where it reads them with to_long, so both views stay separate here. */
inline const std::any& any_ref(const pmt_t& p)
{
    if (!is_any(p)) throw std::runtime_error("pmt: not an any");
    return p->aval;
}
inline bool eqv(const pmt_t& a, const pmt_t& b)
{
    if (!a || !b) return a == b;
    if (a->kind != b->kind) return false;
    if (a->kind == pmt_base::K_SYMBOL) return a->sym == b->sym;
    if (a->kind == pmt_base::K_LONG) return a->lval == b->lval;
    return a == b;
}
inline pmt_t get_PMT_NIL() { return pmt_t(); }
#define PMT_NIL (::pmt::get_PMT_NIL())
}  // namespace pmt
