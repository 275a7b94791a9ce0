// host/Support.h — PointMatcherSupport: the plugin runtime of the reference, re-created without
// Boost / Eigen / yaml-cpp so that it builds in this image.
//
//   Parametrizable  pointmatcher/Parametrizable.{h,cpp}  (string parameters with doc / default /
//                   min / max, bounds checked at construction, "used" tracking)
//   Registrar       pointmatcher/Registrar.h:75-230       (name -> factory, createFromYAML)
//   exceptions      same names and base classes as the reference
//   Yaml            the block-style subset the reference's configuration files use
//                   (doc/Configuration.md, examples/data/*.yaml), replacing contrib/yaml-cpp-pm
#pragma once

#include <cmath>
#include <cstdlib>
#include <istream>
#include <limits>
#include <map>
#include <memory>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace PointMatcherSupport {

// ---- exceptions ----------------------------------------------------------------------------------
struct InvalidModuleType : std::runtime_error {  // PointMatcher.h:83-88
    explicit InvalidModuleType(const std::string& reason) : std::runtime_error(reason) {}
};
struct TransformationError : std::runtime_error {  // PointMatcher.h:89-94
    explicit TransformationError(const std::string& reason) : std::runtime_error(reason) {}
};
struct ConfigurationError : std::runtime_error {  // PointMatcher.h:95-100
    explicit ConfigurationError(const std::string& reason) : std::runtime_error(reason) {}
};
struct InvalidElement : std::runtime_error {  // Registrar.h:69-72
    explicit InvalidElement(const std::string& reason) : std::runtime_error(reason) {}
};

// ---- lexical casts (Parametrizable.h:53-94) --------------------------------------------------------
template <typename Target>
inline Target lexical_cast_scalar_to_string(const std::string& arg) {
    if (arg == "inf") return std::numeric_limits<Target>::infinity();
    if (arg == "-inf") return -std::numeric_limits<Target>::infinity();
    if (arg == "nan") return std::numeric_limits<Target>::quiet_NaN();
    std::istringstream is(arg);
    Target v;
    is >> v;
    if (is.fail() || !(is >> std::ws).eof()) throw std::runtime_error("bad lexical cast: '" + arg + "'");
    return v;
}
template <typename Target>
inline Target lexical_cast(const std::string& arg) {
    std::istringstream is(arg);
    Target v;
    is >> v;
    if (is.fail() || !(is >> std::ws).eof()) throw std::runtime_error("bad lexical cast: '" + arg + "'");
    return v;
}
template <>
inline float lexical_cast<float>(const std::string& arg) { return lexical_cast_scalar_to_string<float>(arg); }
template <>
inline double lexical_cast<double>(const std::string& arg) { return lexical_cast_scalar_to_string<double>(arg); }
template <>
inline std::string lexical_cast<std::string>(const std::string& arg) { return arg; }
template <>
inline bool lexical_cast<bool>(const std::string& arg) { return lexical_cast<int>(arg) != 0; }

template <typename S>
std::string toParam(const S& value) {
    std::ostringstream os;
    os.precision(std::numeric_limits<double>::max_digits10);
    os << value;
    return os.str();
}

// ---- Parametrizable (Parametrizable.h:98-175, Parametrizable.cpp:170-207) ---------------------------
struct Parametrizable {
    struct InvalidParameter : std::runtime_error {
        explicit InvalidParameter(const std::string& reason) : std::runtime_error(reason) {}
    };
    typedef bool (*LexicalComparison)(std::string a, std::string b);
    template <typename S>
    static bool Comp(std::string a, std::string b) { return lexical_cast<S>(a) < lexical_cast<S>(b); }

    struct ParameterDoc {
        std::string name, doc, defaultValue, minValue, maxValue;
        LexicalComparison comp;
        ParameterDoc(const std::string& name, const std::string& doc, const std::string& defaultValue, const std::string& minValue,
                     const std::string& maxValue, LexicalComparison comp)
            : name(name), doc(doc), defaultValue(defaultValue), minValue(minValue), maxValue(maxValue), comp(comp) {}
        ParameterDoc(const std::string& name, const std::string& doc, const std::string& defaultValue)
            : name(name), doc(doc), defaultValue(defaultValue), comp(nullptr) {}
    };
    typedef std::vector<ParameterDoc> ParametersDoc;
    typedef std::string Parameter;
    typedef std::map<std::string, Parameter> Parameters;
    typedef std::set<std::string> ParametersUsed;

    const std::string className;
    const ParametersDoc parametersDoc;
    Parameters parameters;
    ParametersUsed parametersUsed;

    Parametrizable() {}
    Parametrizable(const std::string& className, const ParametersDoc paramsDoc, const Parameters& params)
        : className(className), parametersDoc(paramsDoc) {
        // fill current parameters from either values passed as argument, or default value
        for (const ParameterDoc& d : parametersDoc) {
            const std::string& paramName = d.name;
            Parameters::const_iterator paramIt = params.find(paramName);
            if (paramIt != params.end()) {
                const std::string& val = paramIt->second;
                if (d.comp) {
                    bool tooSmall, tooLarge;
                    try {
                        tooSmall = d.comp(val, d.minValue);
                        tooLarge = d.comp(d.maxValue, val);
                    } catch (const std::runtime_error&) {
                        throw InvalidParameter("Value " + val + " of parameter " + paramName + " in class " + className + " cannot be parsed");
                    }
                    if (tooSmall)
                        throw InvalidParameter("Value " + val + " of parameter " + paramName + " in class " + className +
                                               " is smaller than minimum admissible value " + d.minValue);
                    if (tooLarge)
                        throw InvalidParameter("Value " + val + " of parameter " + paramName + " in class " + className +
                                               " is larger than maximum admissible value " + d.maxValue);
                }
                parameters[paramName] = val;
            } else {
                parameters[paramName] = d.defaultValue;
            }
        }
    }
    virtual ~Parametrizable() {}

    std::string getParamValueString(const std::string& paramName) {
        Parameters::const_iterator paramIt = parameters.find(paramName);
        if (paramIt == parameters.end()) throw InvalidParameter("Parameter " + paramName + " does not exist in class " + className);
        parametersUsed.insert(paramIt->first);
        return paramIt->second;
    }
    template <typename S>
    S get(const std::string& paramName) {
        return lexical_cast<S>(getParamValueString(paramName));
    }
};

// ---- Yaml: block-style subset -----------------------------------------------------------------------
struct YamlNode {
    enum Type { Null, Scalar, Sequence, Map } type = Null;
    std::string scalar;
    std::vector<YamlNode> seq;
    std::vector<std::pair<std::string, YamlNode>> map;  // insertion order kept
    const YamlNode* find(const std::string& key) const {
        for (const auto& kv : map)
            if (kv.first == key) return &kv.second;
        return nullptr;
    }
};

class YamlParser {
public:
    // throws std::runtime_error on malformed input
    static YamlNode parse(std::istream& in) {
        YamlParser p;
        std::string line;
        while (std::getline(in, line)) {
            const size_t hash = findComment(line);
            if (hash != std::string::npos) line.erase(hash);
            while (!line.empty() && (line.back() == ' ' || line.back() == '\t' || line.back() == '\r')) line.pop_back();
            size_t indent = 0;
            while (indent < line.size() && line[indent] == ' ') ++indent;
            if (indent == line.size()) continue;
            if (line[indent] == '\t') throw std::runtime_error("yaml: tabs are not allowed for indentation");
            if (line.compare(indent, 3, "---") == 0 || line.compare(indent, 3, "...") == 0) continue;
            p.lines.push_back({(int)indent, line.substr(indent)});
        }
        size_t pos = 0;
        if (p.lines.empty()) return YamlNode();
        YamlNode root = p.parseBlock(pos, p.lines[0].indent);
        if (pos != p.lines.size()) throw std::runtime_error("yaml: unexpected indentation near '" + p.lines[pos].text + "'");
        return root;
    }

private:
    struct Line { int indent; std::string text; };
    std::vector<Line> lines;

    static size_t findComment(const std::string& s) {
        for (size_t i = 0; i < s.size(); ++i)
            if (s[i] == '#' && (i == 0 || s[i - 1] == ' ' || s[i - 1] == '\t')) return i;
        return std::string::npos;
    }
    static std::string trim(const std::string& s) {
        size_t a = 0, b = s.size();
        while (a < b && (s[a] == ' ' || s[a] == '\t')) ++a;
        while (b > a && (s[b - 1] == ' ' || s[b - 1] == '\t')) --b;
        std::string r = s.substr(a, b - a);
        if (r.size() >= 2 && ((r.front() == '"' && r.back() == '"') || (r.front() == '\'' && r.back() == '\''))) r = r.substr(1, r.size() - 2);
        return r;
    }
    static bool splitKey(const std::string& text, std::string& key, std::string& rest) {
        const size_t c = text.find(':');
        if (c == std::string::npos) return false;
        if (c + 1 < text.size() && text[c + 1] != ' ') return false;
        key = trim(text.substr(0, c));
        rest = trim(text.substr(c + 1));
        return !key.empty();
    }
    YamlNode scalarNode(const std::string& s) {
        YamlNode n;
        n.type = YamlNode::Scalar;
        n.scalar = s;
        return n;
    }
    // a block of lines sharing `indent`: a sequence ("- ...") or a map ("key: ...")
    YamlNode parseBlock(size_t& pos, int indent) {
        YamlNode node;
        if (lines[pos].text.compare(0, 2, "- ") == 0 || lines[pos].text == "-") {
            node.type = YamlNode::Sequence;
            while (pos < lines.size() && lines[pos].indent == indent && (lines[pos].text.compare(0, 2, "- ") == 0 || lines[pos].text == "-")) {
                const std::string item = lines[pos].text.size() > 2 ? trim(lines[pos].text.substr(2)) : std::string();
                const int itemIndent = indent + 2;
                std::string key, rest;
                if (item.empty()) {
                    ++pos;
                    if (pos < lines.size() && lines[pos].indent > indent) node.seq.push_back(parseBlock(pos, lines[pos].indent));
                    else node.seq.push_back(YamlNode());
                } else if (splitKey(item, key, rest)) {
                    // "- Name:" followed by an indented parameter map, or "- key: value" map item
                    lines[pos].indent = itemIndent;
                    lines[pos].text = item;
                    node.seq.push_back(parseBlock(pos, itemIndent));
                } else {
                    node.seq.push_back(scalarNode(item));
                    ++pos;
                }
            }
            return node;
        }
        node.type = YamlNode::Map;
        while (pos < lines.size() && lines[pos].indent == indent) {
            std::string key, rest;
            if (!splitKey(lines[pos].text, key, rest)) {
                if (node.map.empty()) {  // a bare scalar block
                    YamlNode s = scalarNode(trim(lines[pos].text));
                    ++pos;
                    return s;
                }
                throw std::runtime_error("yaml: expected 'key: value' near '" + lines[pos].text + "'");
            }
            ++pos;
            if (!rest.empty()) node.map.push_back({key, scalarNode(rest)});
            else if (pos < lines.size() && lines[pos].indent > indent) node.map.push_back({key, parseBlock(pos, lines[pos].indent)});
            else node.map.push_back({key, YamlNode()});
        }
        if (pos < lines.size() && lines[pos].indent > indent) throw std::runtime_error("yaml: unexpected indentation near '" + lines[pos].text + "'");
        return node;
    }
};

// Registrar.cpp:12-32: a module is either a bare name or {name: {param: value, ...}}
inline void getNameParamsFromYAML(const YamlNode& module, std::string& name, Parametrizable::Parameters& params) {
    if (module.type == YamlNode::Scalar) {
        name = module.scalar;
    } else if (module.type == YamlNode::Map && module.map.size() == 1) {
        name = module.map[0].first;
        const YamlNode& p = module.map[0].second;
        if (p.type == YamlNode::Map) {
            for (const auto& kv : p.map) {
                if (kv.second.type != YamlNode::Scalar) throw InvalidElement("parameter " + kv.first + " of module " + name + " is not a scalar");
                params[kv.first] = kv.second.scalar;
            }
        } else if (p.type != YamlNode::Null) {
            throw InvalidElement("parameters of module " + name + " must be a map");
        }
    } else {
        throw InvalidElement("a module must be a name or a single-entry map");
    }
}

// ---- Registrar (Registrar.h:75-218) --------------------------------------------------------------------
template <typename Interface>
struct Registrar {
    typedef Interface TargetType;
    struct ClassDescriptor {
        virtual ~ClassDescriptor() {}
        virtual std::shared_ptr<Interface> createInstance(const std::string& className, const Parametrizable::Parameters& params) const = 0;
        virtual const std::string description() const = 0;
        virtual const Parametrizable::ParametersDoc availableParameters() const = 0;
    };
    template <typename C>
    struct GenericClassDescriptor : public ClassDescriptor {
        std::shared_ptr<Interface> createInstance(const std::string& className, const Parametrizable::Parameters& params) const override {
            std::shared_ptr<C> instance = std::make_shared<C>(params);
            for (const auto& param : params)
                if (instance->parametersUsed.find(param.first) == instance->parametersUsed.end())
                    throw Parametrizable::InvalidParameter("Parameter " + param.first + " for module " + className + " was set but is not used");
            return instance;
        }
        const std::string description() const override { return C::description(); }
        const Parametrizable::ParametersDoc availableParameters() const override { return C::availableParameters(); }
    };
    template <typename C>
    struct GenericClassDescriptorNoParam : public ClassDescriptor {
        std::shared_ptr<Interface> createInstance(const std::string& className, const Parametrizable::Parameters& params) const override {
            for (const auto& param : params)
                throw Parametrizable::InvalidParameter("Parameter " + param.first + " was set but module " + className + " dos not use any parameter");
            return std::make_shared<C>();
        }
        const std::string description() const override { return C::description(); }
        const Parametrizable::ParametersDoc availableParameters() const override { return Parametrizable::ParametersDoc(); }
    };

protected:
    typedef std::map<std::string, std::shared_ptr<ClassDescriptor>> DescriptorMap;
    DescriptorMap classes;

public:
    void reg(const std::string& name, std::shared_ptr<ClassDescriptor> descriptor) { classes.insert(std::make_pair(name, descriptor)); }
    std::shared_ptr<ClassDescriptor> getDescriptor(const std::string& name) const {
        auto it = classes.find(name);
        if (it == classes.end()) throw InvalidElement("Trying to instanciate unknown element " + name + " from registrar");
        return it->second;
    }
    std::shared_ptr<Interface> create(const std::string& name, const Parametrizable::Parameters& params = Parametrizable::Parameters()) const {
        return getDescriptor(name)->createInstance(name, params);
    }
    std::shared_ptr<Interface> createFromYAML(const YamlNode& module) const {
        std::string name;
        Parametrizable::Parameters params;
        getNameParamsFromYAML(module, name, params);
        return create(name, params);
    }
    const std::string getDescription(const std::string& name) const { return getDescriptor(name)->description(); }
    void dump(std::ostream& stream) const {
        for (const auto& it : classes) stream << "- " << it.first << "\n";
    }
    typename DescriptorMap::const_iterator begin() const { return classes.begin(); }
    typename DescriptorMap::const_iterator end() const { return classes.end(); }
};

#define REG(name) name##Registrar
#define DEF_REGISTRAR(name) PointMatcherSupport::Registrar<name> name##Registrar;
#define DEF_REGISTRAR_IFACE(name, ifaceName) PointMatcherSupport::Registrar<ifaceName> name##Registrar;
#define ADD_TO_REGISTRAR(name, elementName, element)                                                                   \
    {                                                                                                                  \
        typedef typename PointMatcherSupport::Registrar<name>::template GenericClassDescriptor<element> Desc;          \
        name##Registrar.reg(#elementName, std::make_shared<Desc>());                                                   \
    }
#define ADD_TO_REGISTRAR_NO_PARAM(name, elementName, element)                                                          \
    {                                                                                                                  \
        typedef typename PointMatcherSupport::Registrar<name>::template GenericClassDescriptorNoParam<element> Desc;   \
        name##Registrar.reg(#elementName, std::make_shared<Desc>());                                                   \
    }

}  // namespace PointMatcherSupport
