/*
 * oracle/bow_glue.cc — C entry points around the reference's OWN DBoW2 vocabulary code (oracle/_ref build, TEST
 * INFRASTRUCTURE ONLY): Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h is a header template and is compiled verbatim, as
 * ORBVocabulary.h instantiates it (TemplatedVocabulary<FORB::TDescriptor, FORB>): loadFromTextFile (node order, word
 * numbering, weights), transform (tree descent, levelsup, weighting, normalisation flow) and score are the reference's.
 * What the snapshot does NOT contain are DBoW2's .cpp files; their few functions are restated below from DBoW2's
 * published algorithm (FORB::distance / fromString, BowVector::addWeight / addIfNotExist / normalize,
 * FeatureVector::addFeature, the scoring objects) — small leaf functions, each a handful of lines.
 */
#include <cmath>
#include <cstdint>
#include <cstring>
#include <sstream>
#include <string>
#include <vector>
#include <opencv2/core/core.hpp>
#include "Thirdparty/DBoW2/DBoW2/FORB.h"
#include "Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h"

namespace DUtils {
bool Random::m_already_seeded = false;
void Random::SeedRand() {}
void Random::SeedRandOnce() {}
void Random::SeedRand(int) {}
void Random::SeedRandOnce(int) {}
int Random::RandomInt(int min, int) { return min; }                /* only k-means (vocabulary creation) draws numbers */
}

namespace DBoW2 {
const int FORB::L = 32;
void FORB::meanValue(const std::vector<FORB::pDescriptor>&, FORB::TDescriptor&) {}      /* k-means only */
int FORB::distance(const FORB::TDescriptor& a, const FORB::TDescriptor& b)
{
    const int* pa = a.ptr<int32_t>(); const int* pb = b.ptr<int32_t>();
    int dist = 0;
    for (int i = 0; i < 8; i++, pa++, pb++) {
        unsigned int v = *pa ^ *pb;
        v = v - ((v >> 1) & 0x55555555);
        v = (v & 0x33333333) + ((v >> 2) & 0x33333333);
        dist += (((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) >> 24;
    }
    return dist;
}
std::string FORB::toString(const FORB::TDescriptor& a)
{
    std::stringstream ss;
    const unsigned char* p = a.ptr<unsigned char>();
    for (int i = 0; i < a.cols; ++i, ++p) ss << (int)*p << " ";
    return ss.str();
}
void FORB::fromString(FORB::TDescriptor& a, const std::string& s)
{
    a = cv::Mat(1, FORB::L, CV_8U);
    unsigned char* p = a.ptr<unsigned char>();
    std::stringstream ss(s);
    for (int i = 0; i < FORB::L; ++i, ++p) { int n; ss >> n; if (!ss.fail()) *p = (unsigned char)n; }
}
void FORB::toMat32F(const std::vector<TDescriptor>&, cv::Mat&) {}
void FORB::toMat8U(const std::vector<TDescriptor>&, cv::Mat&) {}

BowVector::BowVector(void) {}
BowVector::~BowVector(void) {}
void BowVector::addWeight(WordId id, WordValue v)
{
    BowVector::iterator vit = this->lower_bound(id);
    if (vit != this->end() && !(this->key_comp()(id, vit->first))) vit->second += v;
    else this->insert(vit, BowVector::value_type(id, v));
}
void BowVector::addIfNotExist(WordId id, WordValue v)
{
    BowVector::iterator vit = this->lower_bound(id);
    if (vit == this->end() || (this->key_comp()(id, vit->first))) this->insert(vit, BowVector::value_type(id, v));
}
void BowVector::normalize(LNorm norm_type)
{
    double norm = 0.0;
    BowVector::iterator it;
    if (norm_type == DBoW2::L1) { for (it = begin(); it != end(); ++it) norm += fabs(it->second); }
    else { for (it = begin(); it != end(); ++it) norm += it->second * it->second; norm = sqrt(norm); }
    if (norm > 0.0) for (it = begin(); it != end(); ++it) it->second /= norm;
}
FeatureVector::FeatureVector(void) {}
FeatureVector::~FeatureVector(void) {}
void FeatureVector::addFeature(NodeId id, unsigned int i_feature)
{
    FeatureVector::iterator vit = this->lower_bound(id);
    if (vit != this->end() && vit->first == id) vit->second.push_back(i_feature);
    else { vit = this->insert(vit, FeatureVector::value_type(id, std::vector<unsigned int>())); vit->second.push_back(i_feature); }
}
const double GeneralScoring::LOG_EPS = log(2.220446049250313e-16);
double L1Scoring::score(const BowVector& v1, const BowVector& v2) const
{
    BowVector::const_iterator v1_it = v1.begin(), v2_it = v2.begin();
    const BowVector::const_iterator v1_end = v1.end(), v2_end = v2.end();
    double score = 0;
    while (v1_it != v1_end && v2_it != v2_end) {
        const WordValue& vi = v1_it->second; const WordValue& wi = v2_it->second;
        if (v1_it->first == v2_it->first) { score += fabs(vi - wi) - fabs(vi) - fabs(wi); ++v1_it; ++v2_it; }
        else if (v1_it->first < v2_it->first) v1_it = v1.lower_bound(v2_it->first);
        else v2_it = v2.lower_bound(v1_it->first);
    }
    score = -score / 2.0;
    return score;
}
/* ORBvoc.txt uses L1_NORM; the other scoring objects exist only to satisfy the vtable of createScoringObject() */
double L2Scoring::score(const BowVector&, const BowVector&) const { return 0; }
double ChiSquareScoring::score(const BowVector&, const BowVector&) const { return 0; }
double KLScoring::score(const BowVector&, const BowVector&) const { return 0; }
double BhattacharyyaScoring::score(const BowVector&, const BowVector&) const { return 0; }
double DotProductScoring::score(const BowVector&, const BowVector&) const { return 0; }
}

typedef DBoW2::TemplatedVocabulary<DBoW2::FORB::TDescriptor, DBoW2::FORB> ORBVocabulary;   /* ORBVocabulary.h:31-32 */

extern "C" {
void* bref_vocab_load(const char* text_file)
{
    ORBVocabulary* v = new ORBVocabulary();
    if (!v->loadFromTextFile(text_file)) { delete v; return 0; }
    return v;
}
void bref_vocab_destroy(void* v) { delete static_cast<ORBVocabulary*>(v); }
int bref_vocab_words(void* v) { return (int)static_cast<ORBVocabulary*>(v)->size(); }
/* Frame::ComputeBoW: transform(vCurrentDesc, mBowVec, mFeatVec, levelsup) -> BowVector and FeatureVector flattened */
void bref_vocab_transform(void* vv, const uint8_t* desc, int n, int levelsup, int32_t* bow_id, double* bow_val, int32_t* n_bow,
                          int32_t* fv_node, int32_t* fv_off, int32_t* fv_feat, int32_t* n_fv)
{
    ORBVocabulary* v = static_cast<ORBVocabulary*>(vv);
    std::vector<cv::Mat> feats((size_t)n);
    for (int i = 0; i < n; i++) { feats[i] = cv::Mat(1, 32, CV_8U); memcpy(feats[i].ptr<uint8_t>(), desc + 32 * (size_t)i, 32); }
    DBoW2::BowVector bv; DBoW2::FeatureVector fv;
    v->transform(feats, bv, fv, levelsup);
    int k = 0;
    for (DBoW2::BowVector::const_iterator it = bv.begin(); it != bv.end(); ++it, ++k) { bow_id[k] = (int32_t)it->first; bow_val[k] = it->second; }
    *n_bow = k;
    int a = 0, o = 0;
    for (DBoW2::FeatureVector::const_iterator it = fv.begin(); it != fv.end(); ++it, ++a) {
        fv_node[a] = (int32_t)it->first; fv_off[a] = o;
        for (size_t j = 0; j < it->second.size(); j++) fv_feat[o++] = (int32_t)it->second[j];
    }
    fv_off[a] = o; *n_fv = a;
}
double bref_vocab_score(void* vv, const int32_t* id1, const double* v1, int n1, const int32_t* id2, const double* v2, int n2)
{
    DBoW2::BowVector a, b;
    for (int i = 0; i < n1; i++) a.insert(a.end(), std::make_pair((DBoW2::WordId)id1[i], v1[i]));
    for (int i = 0; i < n2; i++) b.insert(b.end(), std::make_pair((DBoW2::WordId)id2[i], v2[i]));
    return static_cast<ORBVocabulary*>(vv)->score(a, b);
}
}
