// config_io.cpp -- readers / writer for the reference's configuration, contact-matrix and data files, and the
// assembly of one calibration project from a reference-style tree.  See config_io.hpp for the reference functions
// each piece stands in for.
#include "config_io.hpp"

#include <algorithm>
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <fstream>
#include <functional>
#include <sstream>

namespace epidemic {

namespace {

const char* const kBlank = " \t\n\r\f\v";

std::string trimmed(const std::string& s) {
    const size_t b = s.find_first_not_of(kBlank);
    if (b == std::string::npos) return {};
    return s.substr(b, s.find_last_not_of(kBlank) - b + 1);
}

// One data line of a `name value...` file: the leading word and the numbers that follow it.  `junk` is set when a
// token after the numbers is not a number (a comment marker, a stray word): formatted extraction stops there.
struct ConfigLine {
    int number = 0;
    std::string text, name;
    std::vector<double> values;
    bool junk = false;
};

// A number the way `std::istream >> double` accepts it: optional sign, digits with an optional point, optional
// exponent; no "inf", "nan" or hex.  Returns the characters consumed (0 = not a number).
size_t stream_number(const std::string& tok, double& out) {
    size_t i = 0, digits = 0;
    if (i < tok.size() && (tok[i] == '+' || tok[i] == '-')) ++i;
    while (i < tok.size() && std::isdigit(static_cast<unsigned char>(tok[i]))) { ++i; ++digits; }
    if (i < tok.size() && tok[i] == '.') {
        ++i;
        while (i < tok.size() && std::isdigit(static_cast<unsigned char>(tok[i]))) { ++i; ++digits; }
    }
    if (digits == 0) return 0;
    if (i < tok.size() && (tok[i] == 'e' || tok[i] == 'E')) {
        size_t j = i + 1;
        if (j < tok.size() && (tok[j] == '+' || tok[j] == '-')) ++j;
        size_t ed = 0;
        while (j < tok.size() && std::isdigit(static_cast<unsigned char>(tok[j]))) { ++j; ++ed; }
        if (ed > 0) i = j;
    }
    out = std::strtod(tok.substr(0, i).c_str(), nullptr);
    return i;
}

// Calls `fn` for every data line.  `who` names the reader in the FileIOException of an unopenable file.
void for_each_config_line(const std::string& filename, const char* who, const std::string& what,
                          const std::function<void(const ConfigLine&)>& fn) {
    std::ifstream file(filename);
    if (!file.is_open()) throw FileIOException(who, what + filename);
    std::string raw;
    int number = 0;
    while (std::getline(file, raw)) {
        ++number;
        ConfigLine ln;
        ln.number = number;
        ln.text = trimmed(raw);
        if (ln.text.empty() || ln.text[0] == '#') continue;
        std::istringstream words(ln.text);
        std::string tok;
        words >> ln.name;
        while (words >> tok) {
            double v = 0.0;
            const size_t used = stream_number(tok, v);
            if (used > 0) ln.values.push_back(v);
            if (used != tok.size()) { ln.junk = true; break; }
        }
        fn(ln);
    }
}

std::map<std::string, double> read_settings(const std::string& filename, const char* who) {
    std::map<std::string, double> out;
    for_each_config_line(filename, who, "Error opening settings file: ", [&](const ConfigLine& ln) {
        if (ln.values.empty()) throw DataFormatException(who, "Invalid line in settings file: " + ln.text);
        if (ln.values.size() > 1 || ln.junk) throw DataFormatException(who, "Too many values on line in settings file: " + ln.text);
        out[ln.name] = ln.values[0];
    });
    return out;
}

std::vector<std::string> split_cells(const std::string& line) {
    // std::getline(stream, cell, ',') semantics: no quoting, a trailing empty cell is dropped
    std::vector<std::string> cells;
    size_t start = 0;
    while (start < line.size()) {
        const size_t comma = line.find(',', start);
        if (comma == std::string::npos) { cells.push_back(line.substr(start)); return cells; }
        cells.push_back(line.substr(start, comma - start));
        start = comma + 1;
    }
    return cells;
}

VectorXd from_std(const std::vector<double>& v) { return VectorXd::FromPointer(v.data(), static_cast<std::ptrdiff_t>(v.size())); }

std::string sci8(double v) {
    char buf[64];
    std::snprintf(buf, sizeof buf, "%.8e", v);
    return buf;
}
std::string fix1(double v) {
    char buf[64];
    std::snprintf(buf, sizeof buf, "%.1f", v);
    return buf;
}

}  // namespace

// ---- exceptions ---------------------------------------------------------------------------------------------
CSVReadException::CSVReadException(ErrorType type, const std::string& functionName, const std::string& details)
    : DataFormatException(functionName, [&] {
          switch (type) {
              case ErrorType::FileOpenError: return "Could not open file: " + details;
              case ErrorType::InvalidNumberFormat: return "Invalid number format at " + details;
              case ErrorType::NotEnoughRows: return "Not enough rows: " + details;
              case ErrorType::NotEnoughColumns: return "Not enough columns in " + details;
              default: return "Unknown CSV read error: " + details;
          }
      }()),
      type_(type) {}

// ---- configuration files ------------------------------------------------------------------------------------
SEPAIHRDParameters readSEPAIHRDParameters(const std::string& filename, int num_age_classes) {
    const char* who = "readSEPAIHRDParameters";
    SEPAIHRDParameters params;
    struct AgeField { const char* name; VectorXd SEPAIHRDParameters::*member; };
    const AgeField age_fields[] = {{"a", &SEPAIHRDParameters::a}, {"h_infec", &SEPAIHRDParameters::h_infec},
                                   {"p", &SEPAIHRDParameters::p}, {"h", &SEPAIHRDParameters::h},
                                   {"icu", &SEPAIHRDParameters::icu}, {"d_H", &SEPAIHRDParameters::d_H},
                                   {"d_ICU", &SEPAIHRDParameters::d_ICU}, {"d_community", &SEPAIHRDParameters::d_community}};
    for (const AgeField& f : age_fields) params.*(f.member) = VectorXd::Zero(num_age_classes);   // absent lines read as zeros
    struct ScalarField { const char* name; double SEPAIHRDParameters::*member; };
    const ScalarField scalar_fields[] = {
        {"beta", &SEPAIHRDParameters::beta}, {"theta", &SEPAIHRDParameters::theta}, {"sigma", &SEPAIHRDParameters::sigma},
        {"gamma_p", &SEPAIHRDParameters::gamma_p}, {"gamma_A", &SEPAIHRDParameters::gamma_A},
        {"gamma_I", &SEPAIHRDParameters::gamma_I}, {"gamma_H", &SEPAIHRDParameters::gamma_H},
        {"gamma_ICU", &SEPAIHRDParameters::gamma_ICU}, {"E0_multiplier", &SEPAIHRDParameters::E0_multiplier},
        {"P0_multiplier", &SEPAIHRDParameters::P0_multiplier}, {"A0_multiplier", &SEPAIHRDParameters::A0_multiplier},
        {"I0_multiplier", &SEPAIHRDParameters::I0_multiplier}, {"H0_multiplier", &SEPAIHRDParameters::H0_multiplier},
        {"ICU0_multiplier", &SEPAIHRDParameters::ICU0_multiplier}, {"R0_multiplier", &SEPAIHRDParameters::R0_multiplier},
        {"D0_multiplier", &SEPAIHRDParameters::D0_multiplier}, {"runup_days", &SEPAIHRDParameters::runup_days},
        {"seed_exposed", &SEPAIHRDParameters::seed_exposed}};
    std::map<int, double> beta_by_index, kappa_by_index;     // beta_k / kappa_k, 1-based

    auto indexed = [](const std::string& name, const char* prefix, std::map<int, double>& into, double value) {
        const std::string tail = name.substr(std::char_traits<char>::length(prefix));
        try {
            into[std::stoi(tail)] = value;       // "beta_3x" reads as 3, like std::stoi in the reference
        } catch (const std::exception&) {
            // not an index: the reference logs a warning and moves on
        }
    };

    for_each_config_line(filename, who, "Unable to open parameters file: ", [&](const ConfigLine& ln) {
        if (ln.values.empty()) return;                       // "No value found": warning in the reference
        const double scalar = (ln.values.size() == 1) ? ln.values[0] : 0.0;   // several numbers on a scalar line: the reference keeps 0.0
        if (ln.name.rfind("beta_", 0) == 0 && ln.name != "beta_end_times") { indexed(ln.name, "beta_", beta_by_index, scalar); return; }
        if (ln.name.rfind("kappa_", 0) == 0 && ln.name != "kappa_end_times") { indexed(ln.name, "kappa_", kappa_by_index, scalar); return; }
        if (ln.name == "beta_end_times") { params.beta_end_times = ln.values; return; }
        if (ln.name == "kappa_end_times") { params.kappa_end_times = ln.values; return; }
        for (const ScalarField& f : scalar_fields)
            if (ln.name == f.name) {
                params.*(f.member) = scalar;
                return;
            }
        for (const AgeField& f : age_fields)
            if (ln.name == f.name) {
                if (static_cast<int>(ln.values.size()) != num_age_classes)
                    throw DataFormatException(who, "Incorrect number of values for " + ln.name + ". Expected " +
                                                       std::to_string(num_age_classes) + ", got " + std::to_string(ln.values.size()));
                params.*(f.member) = from_std(ln.values);
                return;
            }
        // unrecognised name: warning in the reference
    });

    auto assemble = [](const std::map<int, double>& by_index, std::vector<double>& out) {
        if (by_index.empty()) return;
        out.assign(static_cast<size_t>(std::max(by_index.rbegin()->first, 0)), 0.0);
        for (const auto& [k, v] : by_index)
            if (k >= 1 && static_cast<size_t>(k) <= out.size()) out[static_cast<size_t>(k - 1)] = v;
    };
    assemble(beta_by_index, params.beta_values);
    assemble(kappa_by_index, params.kappa_values);
    return params;
}

std::map<std::string, std::pair<double, double>> readParamBounds(const std::string& filename) {
    const char* who = "readParamBounds";
    std::map<std::string, std::pair<double, double>> bounds;
    for_each_config_line(filename, who, "Error opening param bounds file: ", [&](const ConfigLine& ln) {
        if (ln.values.size() < 2) throw DataFormatException(who, "Invalid line in bounds file: " + ln.text);
        if (ln.values.size() > 2 || ln.junk) throw DataFormatException(who, "Too many values on line in bounds file: " + ln.text);
        bounds[ln.name] = {ln.values[0], ln.values[1]};
    });
    return bounds;
}

std::map<std::string, double> readProposalSigmas(const std::string& filename) {
    const char* who = "readProposalSigmas";
    std::map<std::string, double> sigmas;
    for_each_config_line(filename, who, "Error opening proposal sigmas file: ", [&](const ConfigLine& ln) {
        if (ln.values.empty()) throw DataFormatException(who, "Invalid line in proposal sigmas file: " + ln.text);
        if (ln.values.size() > 1 || ln.junk) throw DataFormatException(who, "Too many values on line in sigmas file: " + ln.text);
        sigmas[ln.name] = ln.values[0];
    });
    return sigmas;
}

std::vector<std::string> readParamsToCalibrate(const std::string& filename) {
    std::vector<std::string> names;     // first word of every data line, in file order; the rest of the line is ignored
    for_each_config_line(filename, "readParamsToCalibrate", "Error opening params_to_calibrate file: ",
                         [&](const ConfigLine& ln) { names.push_back(ln.name); });
    return names;
}

std::map<std::string, double> readMetropolisHastingsSettings(const std::string& f) { return read_settings(f, "readMetropolisHastingsSettings"); }
std::map<std::string, double> readHillClimbingSettings(const std::string& f) { return read_settings(f, "readHillClimbingSettings"); }
std::map<std::string, double> readParticleSwarmSettings(const std::string& f) { return read_settings(f, "readParticleSwarmSettings"); }
std::map<std::string, double> readNUTSSettings(const std::string& f) { return read_settings(f, "readNUTSSettings"); }

void saveCalibrationResults(const std::string& filename, const SEPAIHRDParameters& prm,
                            const std::vector<std::string>& calibrated, double obj_value, const std::string& timestamp_str) {
    std::ofstream file(filename);
    if (!file.is_open()) throw FileIOException("saveCalibrationResults", "Unable to open file for writing: " + filename);
    std::string ts = timestamp_str;
    if (ts.empty()) {
        const std::time_t now = std::time(nullptr);
        char buf[100];
        ts = std::strftime(buf, sizeof buf, "%Y-%m-%d %H:%M:%S", std::localtime(&now)) ? buf : "TIMESTAMP_ERROR";
    }
    auto is_calibrated = [&](const std::string& name) { return std::find(calibrated.begin(), calibrated.end(), name) != calibrated.end(); };
    auto scalar = [&](const std::string& name, double value) {
        file << name << ' ' << sci8(value) << (is_calibrated(name) ? " # [C]" : "") << '\n';
    };
    auto schedule = [&](const char* times_key, const std::vector<double>& times, const char* value_prefix, const std::vector<double>& values) {
        file << times_key;
        for (double t : times) file << ' ' << fix1(t);
        file << '\n';
        for (size_t i = 0; i < values.size(); ++i) scalar(value_prefix + std::to_string(i + 1), values[i]);
    };
    auto per_age = [&](const std::string& base, const VectorXd& values) {
        file << base;
        bool any = false;
        for (std::ptrdiff_t i = 0; i < values.size(); ++i) {
            file << ' ' << sci8(values(i));
            any = any || is_calibrated(base + "_" + std::to_string(i));
        }
        file << (any ? " # [C]" : "") << '\n';
    };

    file << "# Calibrated SEPAIHRD Model Parameters\n"
         << "# Calibration completed: " << ts << '\n'
         << "# Best objective function value: " << sci8(obj_value) << '\n'
         << "# Calibrated parameters are marked with [C] if they were part of the calibration set.\n\n"
         << "# --- Transmission Parameters ---\n";
    schedule("beta_end_times", prm.beta_end_times, "beta_", prm.beta_values);
    scalar("beta", prm.beta);
    scalar("theta", prm.theta);
    file << "\n# --- Disease Progression Rates ---\n";
    scalar("sigma", prm.sigma);
    scalar("gamma_p", prm.gamma_p);
    scalar("gamma_A", prm.gamma_A);
    scalar("gamma_I", prm.gamma_I);
    scalar("gamma_H", prm.gamma_H);
    scalar("gamma_ICU", prm.gamma_ICU);
    file << "\n# --- Age-specific Parameters ---\n";
    per_age("p", prm.p);
    per_age("a", prm.a);
    per_age("h_infec", prm.h_infec);
    per_age("h", prm.h);
    per_age("icu", prm.icu);
    per_age("d_H", prm.d_H);
    per_age("d_ICU", prm.d_ICU);
    if (prm.d_community.size() > 0) per_age("d_community", prm.d_community);
    file << "\n# --- Initial State Multipliers ---\n";
    scalar("E0_multiplier", prm.E0_multiplier);
    scalar("P0_multiplier", prm.P0_multiplier);
    scalar("A0_multiplier", prm.A0_multiplier);
    scalar("I0_multiplier", prm.I0_multiplier);
    scalar("H0_multiplier", prm.H0_multiplier);
    scalar("ICU0_multiplier", prm.ICU0_multiplier);
    scalar("R0_multiplier", prm.R0_multiplier);
    scalar("D0_multiplier", prm.D0_multiplier);
    scalar("runup_days", prm.runup_days);
    scalar("seed_exposed", prm.seed_exposed);
    file << "\n# --- NPI Strategy Parameters ---\n";
    schedule("kappa_end_times", prm.kappa_end_times, "kappa_", prm.kappa_values);
}

// ---- CSV ---------------------------------------------------------------------------------------------------
MatrixXd readMatrixFromCSV(const std::string& filename, int rows, int cols) {
    const std::string who = "epidemic::readMatrixFromCSV";
    using ET = CSVReadException::ErrorType;
    std::ifstream file(filename, std::ios::binary);
    if (!file.is_open()) throw CSVReadException(ET::FileOpenError, who, filename);
    MatrixXd mat(rows, cols);
    std::string line;
    bool have_first = false;
    while (std::getline(file, line))      // the first line that is empty or does not start with "//" is row 1
        if (line.empty() || line.compare(0, 2, "//") != 0) { have_first = true; break; }
    if (!have_first) throw CSVReadException(ET::NotEnoughRows, who, "No data rows found in file: " + filename);
    for (int i = 0; i < rows; ++i) {
        if (i > 0) {
            do {
                if (!std::getline(file, line))
                    throw CSVReadException(ET::NotEnoughRows, who, "expected " + std::to_string(rows) + " rows, found " + std::to_string(i) + " in " + filename);
            } while (line.empty());
        }
        const std::vector<std::string> cells = split_cells(line);
        for (int j = 0; j < cols; ++j) {
            const std::string where = "row " + std::to_string(i + 1);
            if (static_cast<size_t>(j) >= cells.size()) throw CSVReadException(ET::NotEnoughColumns, who, where + " in " + filename);
            try {
                mat(i, j) = std::stod(cells[static_cast<size_t>(j)]);
            } catch (const std::invalid_argument&) {
                throw CSVReadException(ET::InvalidNumberFormat, who, where + ", column " + std::to_string(j + 1) + ": '" + cells[static_cast<size_t>(j)] + "' in " + filename);
            } catch (const std::out_of_range&) {
                throw CSVReadException(ET::InvalidNumberFormat, who, "Number out of range at " + where + ", column " + std::to_string(j + 1) + ": '" + cells[static_cast<size_t>(j)] + "' in " + filename);
            }
        }
    }
    return mat;
}

CalibrationDataFile::CalibrationDataFile(const std::string& filename, const std::string& start_date, const std::string& end_date) {
    const std::string fail = "Failed to initialize CalibrationData from file: " + filename;
    std::ifstream file(filename);
    if (!file.is_open()) throw FileIOException("CalibrationData", fail + " (unable to open)");
    std::string header;
    if (!std::getline(file, header)) throw DataFormatException("CalibrationData", fail + " (empty file)");
    std::map<std::string, int> column;
    {
        const std::vector<std::string> names = split_cells(header);
        for (size_t i = 0; i < names.size(); ++i) column[names[i]] = static_cast<int>(i);
    }
    static const char* const bands[4] = {"0_30", "30_60", "60_80", "80_plus"};
    int needed = 0;
    auto col = [&](const std::string& name) {
        const auto it = column.find(name);
        if (it == column.end()) throw DataFormatException("CalibrationData", "Missing required column: " + name);
        needed = std::max(needed, it->second + 1);
        return it->second;
    };
    auto band_cols = [&](const std::string& prefix) {
        std::vector<int> ix;
        for (const char* b : bands) ix.push_back(col(prefix + "_" + b));
        return ix;
    };
    const int date_col = col("date");
    struct Series { std::vector<int> cols; MatrixXd* into; };
    Series series[] = {{band_cols("new_confirmed"), &new_confirmed_},
                       {band_cols("new_deceased"), &new_deaths_},
                       {band_cols("new_hospitalized_patients"), &new_hosp_},
                       {band_cols("new_intensive_care_patients"), &new_icu_},
                       {band_cols("cumulative_confirmed"), &cum_confirmed_},
                       {band_cols("cumulative_deceased"), &cum_deaths_},
                       {band_cols("cumulative_hospitalized_patients"), &cum_hosp_},
                       {band_cols("cumulative_intensive_care_patients"), &cum_icu_}};
    const std::vector<int> pop_cols = band_cols("population");

    auto in_window = [&](const std::string& date) {
        if (!start_date.empty() && date < start_date) return false;
        if (!end_date.empty() && date > end_date) return false;
        return true;
    };
    auto number = [](const std::string& s) {
        double v = 0.0;
        const auto res = std::from_chars(s.data(), s.data() + s.size(), v);
        if (res.ec != std::errc()) throw DataFormatException("CalibrationData", "Failed to parse value: " + s);
        return v;
    };

    std::vector<std::vector<std::string>> kept;
    std::string line;
    while (std::getline(file, line)) {
        if (line.empty()) continue;
        std::vector<std::string> cells = split_cells(line);
        const std::string date = (static_cast<size_t>(date_col) < cells.size()) ? cells[static_cast<size_t>(date_col)] : std::string();
        if (!in_window(date)) continue;
        if (cells.size() < static_cast<size_t>(needed))
            throw DataFormatException("CalibrationData", fail + " (insufficient columns in data row " + std::to_string(kept.size()) + ")");
        kept.push_back(std::move(cells));
    }
    if (kept.empty()) throw DataFormatException("CalibrationData", fail + " (no data points found in specified date range)");

    const auto rows = static_cast<std::ptrdiff_t>(kept.size());
    for (Series& s : series) s.into->resize(rows, 4);
    population_ = VectorXd(4);
    for (std::ptrdiff_t r = 0; r < rows; ++r) {
        const std::vector<std::string>& cells = kept[static_cast<size_t>(r)];
        dates_.push_back(cells[static_cast<size_t>(date_col)]);
        for (Series& s : series)
            for (int a = 0; a < 4; ++a) (*s.into)(r, a) = number(cells[static_cast<size_t>(s.cols[static_cast<size_t>(a)])]);
    }
    for (int a = 0; a < 4; ++a) population_(a) = number(kept[0][static_cast<size_t>(pop_cols[static_cast<size_t>(a)])]);
}

VectorXd CalibrationDataFile::getInitialSEPAIHRDState(double sigma, double gamma_p, double gamma_a, double gamma_i,
                                                      const VectorXd& p_asymptomatic, const VectorXd& h_hospitalized) const {
    const int n = getNumAgeClasses();
    if (p_asymptomatic.size() != n) throw InvalidParameterException("CalibrationData", "p_asymptomatic vector size mismatch with num_age_classes.");
    if (h_hospitalized.size() != n) throw InvalidParameterException("CalibrationData", "h_hospitalization vector size mismatch with num_age_classes.");
    const VectorXd& N = population_;
    VectorXd state = VectorXd::Zero(SEPAIHRD_NUM_COMPARTMENTS * n);
    auto at = [&](int comp, int age) -> double& { return state(comp * n + age); };
    enum { S, E, P, A, I, H, ICU, R, D, CumH, CumICU };
    for (int i = 0; i < n; ++i) {
        // anchors: the cumulative counts of the first kept day
        double d0 = std::max(cum_deaths_(0, i), 0.0);
        double h0 = std::max(cum_hosp_(0, i), 0.0);
        double u0 = std::max(cum_icu_(0, i), 0.0);
        const double cum_h0 = h0, cum_u0 = u0;
        double i0 = std::max(cum_confirmed_(0, i) - d0, 0.0);
        // unobserved compartments by quasi-steady-state ratios of the progression rates
        const double p_i = std::clamp(p_asymptomatic(i), 0.0, 1.0), q_i = 1.0 - p_i;
        double p0 = (gamma_p > 1e-9 && q_i > 1e-9) ? i0 * gamma_i / (q_i * gamma_p) : i0;
        double a0 = (gamma_a > 1e-9) ? p0 * p_i * gamma_p / gamma_a : p0 * p_i;
        double e0 = (sigma > 1e-9) ? p0 * gamma_p / sigma : p0;
        e0 = std::max(e0, 0.0); p0 = std::max(p0, 0.0); a0 = std::max(a0, 0.0);
        // nobody is counted twice: each observed compartment is capped by what the population has left
        d0 = std::min(d0, N(i));
        u0 = std::min(u0, std::max(0.0, N(i) - d0));
        h0 = std::min(h0, std::max(0.0, N(i) - d0 - u0));
        i0 = std::min(i0, std::max(0.0, N(i) - d0 - u0 - h0));
        const double r0 = std::min(0.0, std::max(0.0, N(i) - d0 - u0 - h0 - i0));
        const double fixed_sum = i0 + h0 + u0 + r0 + d0;
        const double inferred_sum = e0 + p0 + a0;
        const double room = std::max(N(i) - fixed_sum, 0.0);
        if (inferred_sum > room) {
            const double scale = (inferred_sum > 1e-9) ? room / inferred_sum : 0.0;
            e0 *= scale; p0 *= scale; a0 *= scale;
        }
        at(E, i) = e0; at(P, i) = p0; at(A, i) = a0; at(I, i) = i0; at(H, i) = h0; at(ICU, i) = u0; at(R, i) = r0; at(D, i) = d0;
        at(CumH, i) = cum_h0; at(CumICU, i) = cum_u0;
        double non_s = 0.0;
        for (int c = E; c <= D; ++c) non_s += at(c, i);
        at(S, i) = std::max(0.0, N(i) - non_s);
    }
    return state;
}

CalibrationData CalibrationDataFile::toCalibrationData(const VectorXd& initial_state) const {
    return CalibrationData(new_hosp_, new_icu_, new_deaths_, population_, initial_state);
}

// ---- project assembly ---------------------------------------------------------------------------------------
std::shared_ptr<PiecewiseConstantNpiStrategy> createNpiStrategy(const SEPAIHRDParameters& params,
                                                                const std::vector<std::string>& kappa_names,
                                                                const std::map<std::string, std::pair<double, double>>& bounds,
                                                                int fixed_idx) {
    std::map<std::string, std::pair<double, double>> npi_bounds;
    for (const std::string& name : kappa_names) {
        if (name == "kappa_1") continue;       // the fixed baseline has no bounds entry
        const auto it = bounds.find(name);
        if (it != bounds.end()) npi_bounds[name] = it->second;
    }
    const auto first = static_cast<size_t>(fixed_idx);
    const double baseline_kappa = params.kappa_values.at(first);
    const double baseline_end = params.kappa_end_times.at(first);
    std::vector<double> end_times, values;
    std::vector<std::string> names;
    if (params.kappa_end_times.size() > first + 1) {
        end_times.assign(params.kappa_end_times.begin() + fixed_idx + 1, params.kappa_end_times.end());
        values.assign(params.kappa_values.begin() + fixed_idx + 1, params.kappa_values.end());
        names.assign(kappa_names.begin() + fixed_idx + 1, kappa_names.end());
    }
    return std::make_shared<PiecewiseConstantNpiStrategy>(end_times, values, npi_bounds, baseline_kappa, baseline_end, true, names);
}

ReferenceProject loadReferenceProject(const std::string& root, const std::string& start_date, const std::string& end_date, int n) {
    auto path = [&](const char* rel) { return root + (root.empty() || root.back() == '/' ? "" : "/") + rel; };
    ReferenceProject prj;
    prj.data = std::make_shared<CalibrationDataFile>(path("data/processed/processed_data.csv"), start_date, end_date);
    const MatrixXd contacts = readMatrixFromCSV(path("data/contacts.csv"), n, n);
    const VectorXd& N = prj.data->getPopulationByAgeGroup();
    if (N.size() != n) throw DataFormatException("main", "Population data size mismatch");

    prj.params = readSEPAIHRDParameters(path("data/configuration/initial_guess.txt"), n);
    prj.params.N = N;
    prj.params.M_baseline = contacts;
    if (prj.params.kappa_values.size() != prj.params.kappa_end_times.size() || prj.params.beta_values.size() != prj.params.beta_end_times.size())
        throw DataFormatException("main", "Mismatch between end times and values for kappa or beta schedules.");
    std::vector<std::string> kappa_names;
    for (size_t i = 0; i < prj.params.kappa_values.size(); ++i) kappa_names.push_back("kappa_" + std::to_string(i + 1));

    prj.param_bounds = readParamBounds(path("data/configuration/param_bounds.txt"));
    prj.proposal_sigmas = readProposalSigmas(path("data/configuration/proposal_sigmas.txt"));
    prj.params_to_calibrate = readParamsToCalibrate(path("data/configuration/params_to_calibrate.txt"));

    // the grid is fixed from int(runup_days) of the initial file (quirk Q3)
    const int runup = static_cast<int>(prj.params.runup_days);
    const int num_days = prj.data->getNumDataPoints();
    for (int t = -runup; t < num_days; ++t) prj.time_points.push_back(static_cast<double>(t));

    prj.data_initial_state = prj.data->getInitialSEPAIHRDState(prj.params.sigma, prj.params.gamma_p, prj.params.gamma_A,
                                                               prj.params.gamma_I, prj.params.p, prj.params.h);
    prj.initial_state = prj.data_initial_state;
    VectorXd& x = prj.initial_state;
    if (prj.params.runup_days > 0 && prj.params.seed_exposed > 0) {
        const double total = N.sum();
        for (int i = 0; i < n; ++i) {
            x(1 * n + i) = prj.params.seed_exposed * (N(i) / total);
            for (int c = 2; c < SEPAIHRD_NUM_COMPARTMENTS; ++c) x(c * n + i) = 0.0;
        }
    } else {
        const double mult[8] = {prj.params.E0_multiplier, prj.params.P0_multiplier, prj.params.A0_multiplier, prj.params.I0_multiplier,
                                prj.params.H0_multiplier, prj.params.ICU0_multiplier, prj.params.R0_multiplier, prj.params.D0_multiplier};
        for (int c = 1; c <= 8; ++c)
            for (int i = 0; i < n; ++i) x(c * n + i) *= mult[c - 1];
    }
    for (int i = 0; i < n; ++i) {
        double non_s = 0.0;
        for (int c = 1; c <= 8; ++c) non_s += x(c * n + i);
        x(i) = (non_s > N(i)) ? 0.0 : N(i) - non_s;      // main() clamps S to 0 with a warning
    }

    prj.npi_strategy = createNpiStrategy(prj.params, kappa_names, prj.param_bounds, 0);
    prj.model = std::make_shared<AgeSEPAIHRDModel>(prj.params, prj.npi_strategy);
    return prj;
}

}  // namespace epidemic
