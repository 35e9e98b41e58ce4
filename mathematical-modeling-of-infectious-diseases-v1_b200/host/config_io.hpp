// config_io.hpp -- the reference's on-disk formats for the SEPAIHRD hot path, as a standalone C++ host library
// (SURVEY.md section 8f row 4), so that the B200 evaluator can be driven from a reference-style project tree
// without the reference's own sources.
//
// Same function names, argument meaning and error behaviour as the reference functions they stand in for
// (paths relative to the reference repository):
//   include/utils/ReadCalibrationConfiguration.hpp     readSEPAIHRDParameters, readParamBounds, readProposalSigmas,
//       src/utils/ReadCalibrationConfiguration.cpp:51-420  readParamsToCalibrate, read*Settings, saveCalibrationResults
//   include/utils/ReadContactMatrix.hpp                readMatrixFromCSV        (src/utils/ReadContactMatrix.cpp:8-82)
//   include/utils/GetCalibrationData.hpp               CalibrationData CSV constructor + data-derived initial state
//       src/utils/GetCalibrationData.cpp:15-22, 91-98, 107-234, 236-401
//   src/model/main.cpp:81-130, 188-316                 createNpiStrategy, the assembly of one calibration project
//
// The parsing itself is written from the formats, not from the reference's code: one line tokenizer serves all the
// `name value...` files, one cell splitter serves both CSV readers.
#pragma once

#include <map>
#include <memory>
#include <string>
#include <utility>
#include <vector>

#include "epidemic_host.hpp"

namespace epidemic {

// ---- exceptions (include/exceptions/Exceptions.hpp, CSVReadException.hpp) ---------------------------------
class FileIOException : public ModelException { using ModelException::ModelException; };
class DataFormatException : public ModelException { using ModelException::ModelException; };
class CSVReadException : public DataFormatException {
public:
    enum class ErrorType { FileOpenError, InvalidNumberFormat, NotEnoughRows, NotEnoughColumns, Unknown };
    CSVReadException(ErrorType type, const std::string& functionName, const std::string& details);
    ErrorType getErrorType() const noexcept { return type_; }
private:
    ErrorType type_;
};

// ---- `name value...` configuration files --------------------------------------------------------------------
// Lines are trimmed; empty lines and lines starting with '#' are skipped; a trailing "# [C]" marker written by
// saveCalibrationResults ends the numbers of a line (`iss >> value` stops at the first non-number).
SEPAIHRDParameters readSEPAIHRDParameters(const std::string& filename, int num_age_classes);
std::map<std::string, std::pair<double, double>> readParamBounds(const std::string& filename);
std::map<std::string, double> readProposalSigmas(const std::string& filename);
std::vector<std::string> readParamsToCalibrate(const std::string& filename);
std::map<std::string, double> readMetropolisHastingsSettings(const std::string& filename);
std::map<std::string, double> readHillClimbingSettings(const std::string& filename);
std::map<std::string, double> readParticleSwarmSettings(const std::string& filename);
std::map<std::string, double> readNUTSSettings(const std::string& filename);

// Calibrated-parameter writer; the file it writes is a valid input of readSEPAIHRDParameters (round trip tested).
// timestamp_str empty -> local time "%Y-%m-%d %H:%M:%S".
void saveCalibrationResults(const std::string& filename, const SEPAIHRDParameters& parameters,
                            const std::vector<std::string>& actual_calibrated_param_names, double obj_value,
                            const std::string& timestamp_str = "");

// ---- CSV ---------------------------------------------------------------------------------------------------
// Leading "//" comment lines are skipped; then `rows` non-empty lines with at least `cols` comma-separated numbers.
MatrixXd readMatrixFromCSV(const std::string& filename, int rows, int cols);

// The processed daily data (4 fixed age bands 0_30, 30_60, 60_80, 80_plus): rows kept when
// start_date <= date <= end_date by string comparison (empty bound = open), population from the first kept row.
class CalibrationDataFile {
public:
    CalibrationDataFile(const std::string& filename, const std::string& start_date = "", const std::string& end_date = "");
    const std::vector<std::string>& getDates() const { return dates_; }
    const MatrixXd& getNewConfirmedCases() const { return new_confirmed_; }
    const MatrixXd& getNewHospitalizations() const { return new_hosp_; }
    const MatrixXd& getNewICU() const { return new_icu_; }
    const MatrixXd& getNewDeaths() const { return new_deaths_; }
    const MatrixXd& getCumulativeConfirmedCases() const { return cum_confirmed_; }
    const MatrixXd& getCumulativeDeaths() const { return cum_deaths_; }
    const MatrixXd& getCumulativeHospitalizations() const { return cum_hosp_; }
    const MatrixXd& getCumulativeICU() const { return cum_icu_; }
    const VectorXd& getPopulationByAgeGroup() const { return population_; }
    int getNumDataPoints() const { return static_cast<int>(dates_.size()); }
    int getNumAgeClasses() const { return 4; }
    VectorXd getInitialActiveCases() const { return cum_confirmed_.row(0); }
    // quasi-steady-state initial state from the first data row (GetCalibrationData.cpp:107-234); h_hospitalized is
    // only size-checked there, and here.
    VectorXd getInitialSEPAIHRDState(double sigma, double gamma_p, double gamma_a, double gamma_i,
                                     const VectorXd& p_asymptomatic, const VectorXd& h_hospitalized) const;
    // the in-memory object the objective function takes (epidemic_host.hpp)
    CalibrationData toCalibrationData(const VectorXd& initial_state) const;

private:
    std::vector<std::string> dates_;
    MatrixXd new_confirmed_, new_hosp_, new_icu_, new_deaths_, cum_confirmed_, cum_deaths_, cum_hosp_, cum_icu_;
    VectorXd population_;
};

// ---- project assembly (src/model/main.cpp) -----------------------------------------------------------------
// kappa_1 is the fixed baseline; kappa_2.. are the calibratable NPI values with their bounds (main.cpp:81-130).
std::shared_ptr<PiecewiseConstantNpiStrategy> createNpiStrategy(const SEPAIHRDParameters& params,
                                                                const std::vector<std::string>& all_kappa_parameter_names,
                                                                const std::map<std::string, std::pair<double, double>>& overall_param_bounds,
                                                                int fixed_kappa_model_index = 0);

// Everything main() builds before it hands over to the calibrators (main.cpp:188-316): data window, contact matrix,
// parameters, NPI schedule, model, bounds / sigmas / names, the integer time grid -int(runup_days) .. num_days-1,
// the data-derived initial state and the initial state main() itself integrates from (run-up seeding or multipliers).
struct ReferenceProject {
    std::shared_ptr<CalibrationDataFile> data;
    SEPAIHRDParameters params;
    std::shared_ptr<PiecewiseConstantNpiStrategy> npi_strategy;
    std::shared_ptr<AgeSEPAIHRDModel> model;
    std::map<std::string, std::pair<double, double>> param_bounds;
    std::map<std::string, double> proposal_sigmas;
    std::vector<std::string> params_to_calibrate;
    std::vector<double> time_points;
    VectorXd data_initial_state;    // getInitialSEPAIHRDState(...)
    VectorXd initial_state;         // after run-up seeding / multipliers and the S remainder (main.cpp:268-316)
    double abs_error = 1.0e-6, rel_error = 1.0e-6, dt_hint = 1.0;
};
ReferenceProject loadReferenceProject(const std::string& project_root, const std::string& start_date = "2020-03-01",
                                      const std::string& end_date = "2020-12-31", int num_age_classes = 4);

}  // namespace epidemic
