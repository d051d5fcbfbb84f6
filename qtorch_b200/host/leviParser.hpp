// qtorch_b200/host/leviParser.hpp -- ".inp" script reader (">type key value" lines) with the public maps of
// /root/reference/src/leviParser.hpp:26-107.  GPU knobs are ordinary optional keys (e.g. ">int device 0").
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

    bool readInputFile(const std::string &fname) {
        std::ifstream in(fname.c_str());
        if (!in.is_open()) {
            std::cout << "Unable to open file.";
            return false;
        }
        std::string line;
        while (in.good()) {
            std::getline(in, line);
            if (line.empty() || line[0] != '>') continue;
            std::istringstream fields(line);
            std::string kind, key;
            fields >> kind >> key;
            if (kind == ">string") {
                std::string v;
                fields >> v;
                mapString[key] = v;
            } else if (kind == ">bool") {
                std::string v;
                fields >> v;
                if (v == "1" || v == "true" || v == "True" || v == "yes" || v == "Yes") mapBool[key] = true;
                else if (v == "0" || v == "false" || v == "False" || v == "no" || v == "No") mapBool[key] = false;
                else
                    std::cout << "Error in leviParser, " << key
                              << ". bool inputs must be in one of the following forms: 1, true, True, yes, or Yes." << std::endl;
            } else if (kind == ">int") {
                int v = 0;
                fields >> v;
                mapInt[key] = v;
            } else if (kind == ">double") {
                double v = 0.0;
                fields >> v;
                mapDouble[key] = v;
            } else {
                std::cout << "Error. Only the following types are supported in leviParser: string, bool, int, double." << std::endl;
            }
        }
        return true;
    }
};

}  // namespace qtorch
