// qtorch_b200/host/leviParser.hpp -- ".inp" script reader.  A script is a list of lines of the form
//     >string key value      >bool key yes|no|true|false|1|0      >int key 42      >double key 1.5
// (anything that does not start with '>' is a comment).  Same public face as /root/reference/src/leviParser.hpp:26-107
// -- the four maps, the two constructors, readInputFile() -- because main.cpp-style callers index the maps directly.
// GPU knobs are ordinary optional keys (e.g. ">int device 0").
#pragma once
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <string>

namespace qtorch {

class leviParser {
public:
    std::map<std::string, std::string> mapString;
    std::map<std::string, bool> mapBool;
    std::map<std::string, int> mapInt;
    std::map<std::string, double> mapDouble;

    leviParser() {}
    explicit leviParser(const std::string &fname) { readInputFile(fname); }

    // false only when the file cannot be opened; malformed entries are reported on stdout and skipped
    bool readInputFile(const std::string &fname) {
        std::ifstream script(fname.c_str());
        if (!script) {
            std::cout << "Unable to open file.";
            return false;
        }
        for (std::string line; std::getline(script, line);) absorb(line);
        return true;
    }

private:
    // the reference's spellings of a truth value
    static int truthValue(const std::string &word) {
        static const char *const kYes[] = {"1", "true", "True", "yes", "Yes"};
        static const char *const kNo[] = {"0", "false", "False", "no", "No"};
        for (const char *y : kYes) if (word == y) return 1;
        for (const char *n : kNo) if (word == n) return 0;
        return -1;
    }

    template <class T>
    static T next(std::istringstream &fields) {
        T v = T();
        fields >> v;
        return v;
    }

    void absorb(const std::string &line) {
        if (line.empty() || line[0] != '>') return;
        std::istringstream fields(line.substr(1));
        const std::string type = next<std::string>(fields), key = next<std::string>(fields);
        if (type == "int") mapInt[key] = next<int>(fields);
        else if (type == "double") mapDouble[key] = next<double>(fields);
        else if (type == "string") mapString[key] = next<std::string>(fields);
        else if (type == "bool") {
            const int t = truthValue(next<std::string>(fields));
            if (t >= 0) mapBool[key] = (t == 1);
            else std::cout << "Error in leviParser, " << key << ". bool inputs must be in one of the following forms: 1, true, True, yes, or Yes." << std::endl;
        } else {
            std::cout << "Error. Only the following types are supported in leviParser: string, bool, int, double." << std::endl;
        }
    }
};

}  // namespace qtorch
