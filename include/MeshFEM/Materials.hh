// .material files -> Materials::Constant<N> (mirrors Materials.hh / Materials.cc:194-311:
// isotropic[_material], orthotropic[_material], anisotropic / symmetric_material; extra keys such
// as "dim" or "density" are ignored; default material E = 1, nu = 0.3, Materials.hh:408).
#ifndef MESHFEM_B200_MATERIALS_HH
#define MESHFEM_B200_MATERIALS_HH
#include <MeshFEM/ElasticityTensor.hh>
#include <MeshFEM/JSON.hh>

#include <fstream>

namespace Materials {

template <size_t _N>
class Constant {
public:
    typedef ElasticityTensor<Real, _N> ETensor;
    Constant() : m_E(1.0, 0.3) {}
    explicit Constant(const std::string &path) { setFromFile(path); }
    explicit Constant(const ETensor &E) : m_E(E) {}

    void setFromFile(const std::string &materialPath) {
        std::ifstream is(materialPath);
        if (!is.is_open()) throw std::runtime_error("Couldn't open material " + materialPath);
        setFromJson(mjson::json::parse(is));
    }

    void setFromJson(const mjson::json &config) {
        const std::string type = config["type"].str();
        if (type == "isotropic_material" || type == "isotropic") parseIsotropic(config);
        else if (type == "orthotropic_material" || type == "orthotropic") parseOrthotropic(config);
        else if (type == "symmetric_material" || type == "anisotropic") parseAnisotropic(config);
        else throw std::runtime_error("Invalid type.");
    }

    // anisotropic JSON of the current tensor (what getJson of the reference round-trips through)
    mjson::json getJson() const {
        mjson::json j = mjson::json::object();
        j["type"] = mjson::json("anisotropic");
        mjson::json rows = mjson::json::array();
        for (size_t r = 0; r < ETensor::F; ++r) {
            mjson::json row = mjson::json::array();
            for (size_t c = 0; c < ETensor::F; ++c) row.push_back(mjson::json(m_E.D(r, c)));
            rows.push_back(row);
        }
        j["material_matrix"] = rows;
        return j;
    }

    const ETensor &getTensor() const { return m_E; }
    ETensor &getTensor() { return m_E; }
    void setTensor(const ETensor &E) { m_E = E; }
    void setIsotropic(Real E, Real nu) { m_E.setIsotropic(E, nu); }

private:
    ETensor m_E;

    static void parseNVector(size_t n, const mjson::json &j, std::vector<Real> &out) {
        out.clear();
        for (const auto &v : j) out.push_back(v.number());
        if (out.size() != n) throw std::runtime_error("Invalid material parameter vector size");
    }
    void parseIsotropic(const mjson::json &e) { m_E.setIsotropic(e["young"].number(), e["poisson"].number()); }
    void parseOrthotropic(const mjson::json &e) {
        std::vector<Real> poisson, young, shear;
        if (_N == 2) {
            parseNVector(2, e["young"], young);
            parseNVector(2, e["poisson"], poisson);
            parseNVector(1, e["shear"], shear);
            const Real E_x = young[0], E_y = young[1], nu_xy = poisson[0], nu_yx = poisson[1], mu = shear[0];
            m_E.setOrthotropic2D(E_x, E_y, nu_yx, mu);
            if (std::abs(nu_yx / E_y - nu_xy / E_x) > 1e-10) throw std::runtime_error("Orthotopic parameters violate symmetry");
        } else {
            parseNVector(3, e["young"], young);
            parseNVector(6, e["poisson"], poisson);
            parseNVector(3, e["shear"], shear);
            const Real E_x = young[0], E_y = young[1], E_z = young[2];
            const Real nu_yz = poisson[0], nu_zy = poisson[1], nu_zx = poisson[2], nu_xz = poisson[3], nu_xy = poisson[4],
                       nu_yx = poisson[5];
            const Real mu_yz = shear[0], mu_zx = shear[1], mu_xy = shear[2];
            m_E.setOrthotropic3D(E_x, E_y, E_z, nu_yx, nu_zx, nu_zy, mu_yz, mu_zx, mu_xy);
            if ((std::abs(nu_yx / E_y - nu_xy / E_x) > 1e-10) || (std::abs(nu_yz / E_y - nu_zy / E_z) > 1e-10) ||
                (std::abs(nu_zx / E_z - nu_xz / E_x) > 1e-10))
                throw std::runtime_error("Orthotopic parameters violate symmetry");
        }
    }
    void parseAnisotropic(const mjson::json &e) {
        std::runtime_error err("Failed to parse material_matrix");
        size_t row = 0;
        for (const auto &rpt : e["material_matrix"]) {
            if (rpt.size() != flatLen(_N)) throw err;
            size_t col = 0;
            for (const auto &val : rpt) {
                if (row <= col) m_E.D(row, col) = val.number();
                else if (std::abs(m_E.D(row, col) - val.number()) > 1e-10) throw std::runtime_error("Asymmetric material_matrix");
                ++col;
            }
            ++row;
        }
        m_E.symmetrizeFromFull();
    }
};

}  // namespace Materials
#endif
