// json.h — a small recursive-descent JSON reader, enough for the .xalm header (objects, arrays, strings, numbers,
// true/false/null).  The reference vendors nlohmann/json for this (3rdparty/json.hpp); the header is tiny and
// well-formed, so a dependency-free parser keeps the host side self-contained.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace xjson {

struct Value;
using Object = std::vector<std::pair<std::string, Value>>; // insertion-ordered

struct Value {
	enum Kind { Null, Bool, Number, String, Array, Obj } kind = Null;
	bool b = false;
	double num = 0;
	bool is_int = false;
	long long inum = 0;
	unsigned long long unum = 0; // the same digits read as unsigned (xxh3 hashes exceed LLONG_MAX half of the time)
	std::string str;
	std::vector<Value> arr;
	Object obj;

	const Value* find(const std::string& key) const {
		for (auto& kv : obj)
			if (kv.first == key) return &kv.second;
		return nullptr;
	}
	const Value& at(const std::string& key) const {
		const Value* v = find(key);
		if (!v) throw std::out_of_range("json: missing key '" + key + "'");
		return *v;
	}
	bool contains(const std::string& key) const { return find(key) != nullptr; }
};

class Parser {
	const char* p;
	const char* end;

	[[noreturn]] void fail(const char* what) const { throw std::invalid_argument(std::string("json: ") + what); }
	void ws() {
		while (p < end && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) p++;
	}
	static void put_utf8(std::string& s, unsigned cp) {
		if (cp < 0x80) s += (char) cp;
		else if (cp < 0x800) { s += (char) (0xC0 | (cp >> 6)); s += (char) (0x80 | (cp & 0x3F)); }
		else if (cp < 0x10000) { s += (char) (0xE0 | (cp >> 12)); s += (char) (0x80 | ((cp >> 6) & 0x3F)); s += (char) (0x80 | (cp & 0x3F)); }
		else { s += (char) (0xF0 | (cp >> 18)); s += (char) (0x80 | ((cp >> 12) & 0x3F)); s += (char) (0x80 | ((cp >> 6) & 0x3F)); s += (char) (0x80 | (cp & 0x3F)); }
	}
	unsigned hex4() {
		if (end - p < 4) fail("bad \\u escape");
		unsigned v = 0;
		for (int i = 0; i < 4; i++) {
			const char c = *p++;
			v <<= 4;
			if (c >= '0' && c <= '9') v |= c - '0';
			else if (c >= 'a' && c <= 'f') v |= c - 'a' + 10;
			else if (c >= 'A' && c <= 'F') v |= c - 'A' + 10;
			else fail("bad \\u escape");
		}
		return v;
	}
	std::string string() {
		if (*p != '"') fail("expected string");
		p++;
		std::string s;
		while (p < end && *p != '"') {
			if (*p == '\\') {
				if (++p >= end) fail("bad escape");
				switch (*p++) {
					case '"': s += '"'; break;
					case '\\': s += '\\'; break;
					case '/': s += '/'; break;
					case 'b': s += '\b'; break;
					case 'f': s += '\f'; break;
					case 'n': s += '\n'; break;
					case 'r': s += '\r'; break;
					case 't': s += '\t'; break;
					case 'u': {
						unsigned cp = hex4();
						if (cp >= 0xD800 && cp < 0xDC00 && end - p >= 6 && p[0] == '\\' && p[1] == 'u') {
							p += 2;
							const unsigned lo = hex4();
							cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
						}
						put_utf8(s, cp);
						break;
					}
					default: fail("bad escape");
				}
			} else {
				s += *p++;
			}
		}
		if (p >= end) fail("unterminated string");
		p++;
		return s;
	}
	Value value() {
		ws();
		if (p >= end) fail("unexpected end");
		Value v;
		if (*p == '{') {
			v.kind = Value::Obj;
			p++;
			ws();
			if (p < end && *p == '}') { p++; return v; }
			for (;;) {
				ws();
				std::string k = string();
				ws();
				if (p >= end || *p != ':') fail("expected ':'");
				p++;
				v.obj.emplace_back(std::move(k), value());
				ws();
				if (p < end && *p == ',') { p++; continue; }
				if (p < end && *p == '}') { p++; break; }
				fail("expected ',' or '}'");
			}
		} else if (*p == '[') {
			v.kind = Value::Array;
			p++;
			ws();
			if (p < end && *p == ']') { p++; return v; }
			for (;;) {
				v.arr.push_back(value());
				ws();
				if (p < end && *p == ',') { p++; continue; }
				if (p < end && *p == ']') { p++; break; }
				fail("expected ',' or ']'");
			}
		} else if (*p == '"') {
			v.kind = Value::String;
			v.str = string();
		} else if (end - p >= 4 && std::string(p, 4) == "true") { v.kind = Value::Bool; v.b = true; p += 4; }
		else if (end - p >= 5 && std::string(p, 5) == "false") { v.kind = Value::Bool; p += 5; }
		else if (end - p >= 4 && std::string(p, 4) == "null") { p += 4; }
		else {
			const char* s = p;
			bool integral = true;
			if (p < end && (*p == '-' || *p == '+')) p++;
			while (p < end && ((*p >= '0' && *p <= '9') || *p == '.' || *p == 'e' || *p == 'E' || *p == '-' || *p == '+')) {
				if (*p == '.' || *p == 'e' || *p == 'E') integral = false;
				p++;
			}
			if (p == s) fail("unexpected character");
			const std::string t(s, p);
			v.kind = Value::Number;
			v.num = std::strtod(t.c_str(), nullptr);
			v.is_int = integral;
			if (integral) {
				v.inum = std::strtoll(t.c_str(), nullptr, 10);
				v.unum = t[0] == '-' ? (unsigned long long) v.inum : std::strtoull(t.c_str(), nullptr, 10);
			}
		}
		return v;
	}

public:
	static Value parse(const char* data, size_t n) {
		Parser ps;
		ps.p = data;
		ps.end = data + n;
		Value v = ps.value();
		return v;
	}
};

} // namespace xjson
