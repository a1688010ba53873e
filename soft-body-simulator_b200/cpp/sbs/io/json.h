// sbs/io/json.h — the small part of a JSON reader io::load_scene needs (the reference uses
// nlohmann::json, a dependency this library does not take): objects, arrays, strings with the standard
// escapes, numbers, true / false / null.  Lookup of a missing key yields null; iterating null yields
// nothing; converting null or a value of the wrong kind throws json_error.
#ifndef SBS_IO_JSON_H
#define SBS_IO_JSON_H

#include <cstdlib>
#include <istream>
#include <iterator>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace sbs {
namespace io {

struct json_error : std::runtime_error
{
    using std::runtime_error::runtime_error;
};

class json_value
{
  public:
    enum class kind_t { null, boolean, number, string, array, object };
    kind_t kind = kind_t::null;
    bool boolean = false;
    double number = 0.;
    std::string string;
    std::vector<json_value> items;                            // array elements, or object values ...
    std::vector<std::string> keys;                            // ... with their keys, in file order

    bool is_null() const { return kind == kind_t::null; }
    bool contains(std::string const& key) const { return find(key) != nullptr; }
    json_value const& operator[](std::string const& key) const
    {
        static json_value const none{};
        json_value const* v = find(key);
        return v ? *v : none;
    }
    // elements of an array, values of an object (as nlohmann's range-for does), nothing for null
    std::vector<json_value>::const_iterator begin() const { return items.begin(); }
    std::vector<json_value>::const_iterator end() const { return items.end(); }

    template <typename T>
    T get() const
    {
        if constexpr (std::is_same_v<T, std::string>)
        {
            if (kind != kind_t::string)
                throw json_error("json: not a string");
            return string;
        }
        else if constexpr (std::is_same_v<T, bool>)
        {
            if (kind != kind_t::boolean)
                throw json_error("json: not a boolean");
            return boolean;
        }
        else
        {
            if (kind != kind_t::number)
                throw json_error("json: not a number");
            return static_cast<T>(number);
        }
    }

  private:
    json_value const* find(std::string const& key) const
    {
        if (kind != kind_t::object)
            return nullptr;
        for (std::size_t i = 0; i < keys.size(); ++i)
            if (keys[i] == key)
                return &items[i];
        return nullptr;
    }
};

namespace detail {

class json_parser
{
  public:
    explicit json_parser(std::string text) : s_(std::move(text)) {}
    json_value parse()
    {
        json_value v = value();
        skip();
        if (at_ != s_.size())
            fail("trailing characters");
        return v;
    }

  private:
    [[noreturn]] void fail(char const* what) const
    {
        throw json_error("json: " + std::string(what) + " at offset " + std::to_string(at_));
    }
    void skip()
    {
        while (at_ < s_.size() && (s_[at_] == ' ' || s_[at_] == '\t' || s_[at_] == '\n' || s_[at_] == '\r'))
            ++at_;
    }
    bool eat(char c)
    {
        skip();
        if (at_ < s_.size() && s_[at_] == c)
        {
            ++at_;
            return true;
        }
        return false;
    }
    bool word(char const* w)
    {
        std::size_t const n = std::char_traits<char>::length(w);
        if (s_.compare(at_, n, w) != 0)
            return false;
        at_ += n;
        return true;
    }
    static void utf8(std::string& out, unsigned cp)
    {
        if (cp < 0x80)
            out += static_cast<char>(cp);
        else if (cp < 0x800)
        {
            out += static_cast<char>(0xc0 | (cp >> 6));
            out += static_cast<char>(0x80 | (cp & 0x3f));
        }
        else
        {
            out += static_cast<char>(0xe0 | (cp >> 12));
            out += static_cast<char>(0x80 | ((cp >> 6) & 0x3f));
            out += static_cast<char>(0x80 | (cp & 0x3f));
        }
    }
    std::string text()
    {
        std::string out;
        for (;;)
        {
            if (at_ >= s_.size())
                fail("unterminated string");
            char const c = s_[at_++];
            if (c == '"')
                return out;
            if (c != '\\')
            {
                out += c;
                continue;
            }
            if (at_ >= s_.size())
                fail("unterminated escape");
            char const e = s_[at_++];
            switch (e)
            {
            case '"': out += '"'; break;
            case '\\': out += '\\'; break;
            case '/': out += '/'; break;
            case 'b': out += '\b'; break;
            case 'f': out += '\f'; break;
            case 'n': out += '\n'; break;
            case 'r': out += '\r'; break;
            case 't': out += '\t'; break;
            case 'u':
            {
                if (at_ + 4 > s_.size())
                    fail("short \\u escape");
                utf8(out, static_cast<unsigned>(std::strtoul(s_.substr(at_, 4).c_str(), nullptr, 16)));
                at_ += 4;
                break;
            }
            default: fail("bad escape");
            }
        }
    }
    json_value value()
    {
        skip();
        if (at_ >= s_.size())
            fail("unexpected end");
        json_value v;
        char const c = s_[at_];
        if (c == '{')
        {
            ++at_;
            v.kind = json_value::kind_t::object;
            if (eat('}'))
                return v;
            do
            {
                if (!eat('"'))
                    fail("expected a key");
                v.keys.push_back(text());
                if (!eat(':'))
                    fail("expected ':'");
                v.items.push_back(value());
            } while (eat(','));
            if (!eat('}'))
                fail("expected '}'");
        }
        else if (c == '[')
        {
            ++at_;
            v.kind = json_value::kind_t::array;
            if (eat(']'))
                return v;
            do
                v.items.push_back(value());
            while (eat(','));
            if (!eat(']'))
                fail("expected ']'");
        }
        else if (c == '"')
        {
            ++at_;
            v.kind   = json_value::kind_t::string;
            v.string = text();
        }
        else if (word("true"))
        {
            v.kind    = json_value::kind_t::boolean;
            v.boolean = true;
        }
        else if (word("false"))
            v.kind = json_value::kind_t::boolean;
        else if (word("null"))
            v.kind = json_value::kind_t::null;
        else
        {
            char const* first = s_.c_str() + at_;
            char* last        = nullptr;
            v.number          = std::strtod(first, &last);
            if (last == first)
                fail("unexpected character");
            v.kind = json_value::kind_t::number;
            at_ += static_cast<std::size_t>(last - first);
        }
        return v;
    }
    std::string s_;
    std::size_t at_ = 0;
};

} // namespace detail

inline json_value parse_json(std::istream& is)
{
    std::string text((std::istreambuf_iterator<char>(is)), std::istreambuf_iterator<char>());
    return detail::json_parser(std::move(text)).parse();
}

} // namespace io
} // namespace sbs

#endif // SBS_IO_JSON_H
