/* TEST INFRASTRUCTURE ONLY (oracle shim) - stands in for <boost/circular_buffer.hpp>: fixed-capacity ring
 * that overwrites its oldest element, index 0 = oldest.  Only the members the reference uses. */
#pragma once
#include <cstddef>
#include <iterator>
#include <vector>
namespace boost
{
template <typename T>
class circular_buffer
{
public:
    typedef size_t size_type;
    circular_buffer() = default;
    explicit circular_buffer(size_type cap) { set_capacity(cap); }
    void set_capacity(size_type cap)
    {
        std::vector<T> keep;
        const size_type n = d_size < cap ? d_size : cap;
        for (size_type i = d_size - n; i < d_size; i++) keep.push_back((*this)[i]);
        d_buf.assign(cap, T());
        d_head = 0;
        d_size = 0;
        for (auto& v : keep) push_back(v);
    }
    size_type capacity() const { return d_buf.size(); }
    size_type size() const { return d_size; }
    bool full() const { return d_size == d_buf.size(); }
    bool empty() const { return d_size == 0; }
    void clear()
    {
        d_head = 0;
        d_size = 0;
    }
    void push_back(const T& v)
    {
        if (d_buf.empty()) return;
        if (d_size < d_buf.size())
            {
                d_buf[(d_head + d_size) % d_buf.size()] = v;
                d_size++;
            }
        else
            {
                d_buf[d_head] = v;
                d_head = (d_head + 1) % d_buf.size();
            }
    }
    T& operator[](size_type i) { return d_buf[(d_head + i) % d_buf.size()]; }
    const T& operator[](size_type i) const { return d_buf[(d_head + i) % d_buf.size()]; }
    T& front() { return (*this)[0]; }
    T& back() { return (*this)[d_size - 1]; }

    class const_iterator
    {
    public:
        typedef std::forward_iterator_tag iterator_category;
        typedef T value_type;
        typedef std::ptrdiff_t difference_type;
        typedef const T* pointer;
        typedef const T& reference;
        const_iterator(const circular_buffer* b, size_type i) : d_b(b), d_i(i) {}
        reference operator*() const { return (*d_b)[d_i]; }
        pointer operator->() const { return &(*d_b)[d_i]; }
        const_iterator& operator++()
        {
            ++d_i;
            return *this;
        }
        const_iterator operator++(int)
        {
            const_iterator t = *this;
            ++d_i;
            return t;
        }
        bool operator==(const const_iterator& o) const { return d_i == o.d_i; }
        bool operator!=(const const_iterator& o) const { return d_i != o.d_i; }

    private:
        const circular_buffer* d_b;
        size_type d_i;
    };
    typedef const_iterator iterator;
    const_iterator begin() const { return const_iterator(this, 0); }
    const_iterator end() const { return const_iterator(this, d_size); }

private:
    std::vector<T> d_buf;
    size_type d_head{0};
    size_type d_size{0};
};
}  // namespace boost
