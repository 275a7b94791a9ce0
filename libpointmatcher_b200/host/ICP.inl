// host/ICP.inl — ICPChainBase / ICP (PointMatcher.h:652-764, ICP.cpp), included inside
// `struct PointMatcher<T>`.
//
// The 4x4 bookkeeping of ICP::compute stays on the host exactly as in the reference (centre the
// reference on its mean, move the reading into that frame, compose the result).  The iteration
// loop (ICP.cpp:371-430) runs on the device: when the chain consists of the GPU modules and the
// checkers are Counter / Differential, the whole loop is one pmgpu_icp_run call (no host round
// trip per iteration); otherwise every iteration goes through the modules' virtual interface
// with the host checkers deciding, and the device still holds reading, matches and weights.
struct ICPChainBase {
    DataPointsFilters readingDataPointsFilters, readingStepDataPointsFilters, referenceDataPointsFilters;
    Transformations transformations;
    std::shared_ptr<Matcher> matcher;
    OutlierFilters outlierFilters;
    std::shared_ptr<ErrorMinimizer> errorMinimizer;
    TransformationCheckers transformationCheckers;
    std::shared_ptr<Inspector> inspector;
    std::shared_ptr<Logger> logger;

    virtual ~ICPChainBase() {}

    // ICP.cpp:99-113.  The default pre-filters (RandomSampling on the reading, SamplingSurfaceNormal
    // on the reference) run on the host, once per cloud, like the reference's.
    virtual void setDefault() {
        cleanup();
        this->transformations.push_back(std::make_shared<RigidTransformation>());
        this->readingDataPointsFilters.push_back(std::make_shared<RandomSamplingDataPointsFilter>());          // ICP.cpp:105
        this->referenceDataPointsFilters.push_back(std::make_shared<SamplingSurfaceNormalDataPointsFilter>());  // ICP.cpp:106
        this->matcher = std::make_shared<KDTreeMatcher>();
        this->outlierFilters.push_back(std::make_shared<TrimmedDistOutlierFilter>());
        this->errorMinimizer = std::make_shared<PointToPlaneErrorMinimizer>();
        this->transformationCheckers.push_back(std::make_shared<CounterTransformationChecker>());
        this->transformationCheckers.push_back(std::make_shared<DifferentialTransformationChecker>());
        this->inspector = std::make_shared<NullInspector>();
        this->logger = std::make_shared<NullLogger>();
    }

    // ICP.cpp:116-167
    void loadFromYaml(std::istream& in) {
        cleanup();
        PointMatcherSupport::YamlNode doc;
        try {
            doc = PointMatcherSupport::YamlParser::parse(in);
        } catch (const std::runtime_error& e) {
            throw ConfigurationError(std::string("cannot parse the configuration: ") + e.what());
        }
        const PointMatcher& pm = PointMatcher::get();
        std::set<std::string> used;
        auto modules = [&](const char* key, auto& registrar, auto& chain) {
            used.insert(key);
            const PointMatcherSupport::YamlNode* n = doc.find(key);
            if (!n) return;
            if (n->type == PointMatcherSupport::YamlNode::Sequence)
                for (const auto& m : n->seq) chain.push_back(registrar.createFromYAML(m));
            else if (n->type != PointMatcherSupport::YamlNode::Null)
                chain.push_back(registrar.createFromYAML(*n));
        };
        auto module = [&](const char* key, auto& registrar, auto& slot) {
            used.insert(key);
            const PointMatcherSupport::YamlNode* n = doc.find(key);
            if (n && n->type != PointMatcherSupport::YamlNode::Null) slot = registrar.createFromYAML(*n);
        };
        module("logger", pm.LoggerRegistrar, this->logger);
        modules("readingDataPointsFilters", pm.DataPointsFilterRegistrar, this->readingDataPointsFilters);
        modules("readingStepDataPointsFilters", pm.DataPointsFilterRegistrar, this->readingStepDataPointsFilters);
        modules("referenceDataPointsFilters", pm.DataPointsFilterRegistrar, this->referenceDataPointsFilters);
        module("matcher", pm.MatcherRegistrar, this->matcher);
        modules("outlierFilters", pm.OutlierFilterRegistrar, this->outlierFilters);
        module("errorMinimizer", pm.ErrorMinimizerRegistrar, this->errorMinimizer);
        // ICP.cpp:144-148: rigid unless the minimiser is the similarity one (not on the GPU path)
        this->transformations.push_back(std::make_shared<RigidTransformation>());
        modules("transformationCheckers", pm.TransformationCheckerRegistrar, this->transformationCheckers);
        module("inspector", pm.InspectorRegistrar, this->inspector);
        if (!this->inspector) this->inspector = std::make_shared<NullInspector>();
        if (!this->logger) this->logger = std::make_shared<NullLogger>();
        if (doc.type == PointMatcherSupport::YamlNode::Map) {
            for (const auto& kv : doc.map) {
                if (!used.count(kv.first)) throw InvalidModuleType("Module type " + kv.first + " does not exist");  // ICP.cpp:157-166
            }
        } else if (doc.type != PointMatcherSupport::YamlNode::Null) {
            throw ConfigurationError("the configuration must be a map of module types");
        }
    }

    unsigned getPrefilteredReadingPtsCount() const { return prefilteredReadingPtsCount; }
    unsigned getPrefilteredReferencePtsCount() const { return prefilteredReferencePtsCount; }
    bool getMaxNumIterationsReached() const { return maxNumIterationsReached; }

protected:
    unsigned prefilteredReadingPtsCount = 0, prefilteredReferencePtsCount = 0;
    bool maxNumIterationsReached = false;

    void cleanup() {
        transformations.clear();
        readingDataPointsFilters.clear();
        readingStepDataPointsFilters.clear();
        referenceDataPointsFilters.clear();
        matcher.reset();
        outlierFilters.clear();
        errorMinimizer.reset();
        transformationCheckers.clear();
        inspector.reset();
        logger.reset();
    }
};

struct ICP : ICPChainBase {
    TransformationParameters operator()(const DataPoints& readingIn, const DataPoints& referenceIn) {
        const int dim = referenceIn.features.rows();
        return this->compute(readingIn, referenceIn, Matrix::Identity(dim, dim));
    }
    TransformationParameters operator()(const DataPoints& readingIn, const DataPoints& referenceIn, const TransformationParameters& initialTransformationParameters) {
        return this->compute(readingIn, referenceIn, initialTransformationParameters);
    }
    const DataPoints& getReadingFiltered() const { return readingFiltered; }
    size_t getIterationCount() const { return iterationCount; }
    // true when the last compute() ran as one fused device loop (no per-iteration host round trip)
    bool usedFusedLoop() const { return fusedLoop; }

    // ICP.cpp:264-313
    TransformationParameters compute(const DataPoints& readingIn, const DataPoints& referenceIn, const TransformationParameters& T_refIn_dataIn) {
        if (!this->matcher) throw std::runtime_error("You must setup a matcher before running ICP");
        if (!this->errorMinimizer) throw std::runtime_error("You must setup an error minimizer before running ICP");
        if (!this->inspector) throw std::runtime_error("You must setup an inspector before running ICP");
        requireFloat3D(referenceIn.features.rows(), "ICP");
        bindPipeline();
        this->inspector->init();
        const int dim = referenceIn.features.rows();

        // A trailing SurfaceNormalDataPointsFilter that only adds normals is run on the structure the matcher needs anyway
        // (pmgpu_ref_set -> pmgpu_ref_compute_normals -> pmgpu_ref_center: one upload, one build).  Same kernel on the same
        // coordinates as the filter's own run, and normals do not change under the centring translation.
        SurfaceNormalDataPointsFilter* trailingNormals = nullptr;
        if (!this->referenceDataPointsFilters.empty()) {
            auto* last = dynamic_cast<SurfaceNormalDataPointsFilter*>(this->referenceDataPointsFilters.back().get());
            auto* plainMatcher = dynamic_cast<KDTreeMatcher*>(this->matcher.get());
            if (last && plainMatcher && last->keepNormals && !last->keepDensities && !last->keepEigenValues && !last->keepEigenVectors &&
                !last->keepMatchedIds && !last->keepMeanDist && !last->sortEigen && !last->smoothNormals)
                trailingNormals = last;
        }
        // inputs are never mutated (ICP.cpp:285); without reference filters no host copy is needed
        DataPoints& filtered = referenceFilteredHost;  // a member: ErrorElements may be asked for after this call returns
        filtered = DataPoints();
        const size_t hostFilters = this->referenceDataPointsFilters.size() - (trailingNormals ? 1 : 0);
        if (hostFilters > 0) {
            filtered = referenceIn;
            this->referenceDataPointsFilters.init();
            for (size_t i = 0; i < hostFilters; ++i) this->referenceDataPointsFilters[i]->inPlaceFilter(filtered);
        }
        const DataPoints& reference = hostFilters > 0 ? filtered : referenceIn;

        // intermediate frame at the centre of mass of the reference (ICP.cpp:291-299): the mean is
        // the float row sum over all columns divided by N; the matcher is initialised with the
        // centred cloud (ICP.cpp:302) — fused in pmgpu_ref_set_centered for the GPU matcher
        auto* gpuMatcher = dynamic_cast<KDTreeMatcher*>(this->matcher.get());
        if (!gpuMatcher) throw ConfigurationError("ICP: GPU build: the matcher must be the GPU KDTreeMatcher (there is no CPU path)");
        float mean4[4];
        if (trailingNormals) {
            GpuPipeline& g = *pipeline;
            const float* feat = reinterpret_cast<const float*>(reference.features.data());
            g.check(pmgpu_ref_set(g.ctx, feat, reference.features.rows(), reference.features.cols(), nullptr, 0));
            g.check(pmgpu_ref_compute_normals(g.ctx, (int)trailingNormals->knn, (float)trailingNormals->epsilon, (float)trailingNormals->maxDist, 0));
            g.check(pmgpu_ref_center(g.ctx, feat, reference.features.rows(), reference.features.cols(), mean4));
            g.readingKey = nullptr;
        } else {
            gpuMatcher->initCentered(reference, mean4);
        }
        TransformationParameters T_refIn_refMean = Matrix::Identity(dim, dim), T_refMean_refIn = Matrix::Identity(dim, dim);
        for (int r = 0; r < dim - 1; ++r) {
            T_refIn_refMean(r, dim - 1) = T(mean4[r]);
            T_refMean_refIn(r, dim - 1) = T(-mean4[r]);
        }
        this->prefilteredReferencePtsCount = reference.features.cols();
        return computeWithTransformedReference(readingIn, reference, T_refIn_refMean, T_refMean_refIn, T_refIn_dataIn);
    }

protected:
    DataPoints readingFiltered;
    DataPoints referenceFilteredHost;
    size_t iterationCount = 0;
    bool fusedLoop = false;
    std::shared_ptr<GpuPipeline> pipeline;

    void bindPipeline() {
        if (!pipeline) pipeline = std::make_shared<GpuPipeline>(0);
        auto bind = [&](auto* m) {
            if (auto* g = dynamic_cast<GpuBound*>(m)) g->bind(pipeline);
        };
        bind(this->matcher.get());
        bind(this->errorMinimizer.get());
        this->outlierFilters.bind(pipeline);
        for (auto& f : this->outlierFilters) bind(f.get());
        for (auto& f : this->referenceDataPointsFilters) bind(f.get());
        for (auto& f : this->readingDataPointsFilters) bind(f.get());
    }

    // ICP.cpp:316-449
    TransformationParameters computeWithTransformedReference(const DataPoints& readingIn, const DataPoints& reference, const TransformationParameters& T_refIn_refMean,
                                                             const TransformationParameters& T_refMean_refIn, const TransformationParameters& T_refIn_dataIn) {
        const int dim = reference.features.rows();
        if (T_refIn_dataIn.cols() != T_refIn_dataIn.rows()) throw std::runtime_error("The initial transformation matrix must be squared.");
        if (dim != T_refIn_dataIn.cols())
            throw std::runtime_error("The shape of initial transformation matrix must be NxN. Where N is the number of rows in the read/reference scans.");

        // inputs are never mutated (ICP.cpp:324-326): with reading filters they work on a copy; without, the caller's matrix is
        // uploaded as it is (page-locked if the caller registered it, pmgpu_host_pin) and the copy the reference keeps as
        // readingFiltered (ICP.h) is made on a second host thread while the device runs the loop
        DataPoints filteredStorage;
        if (!this->readingDataPointsFilters.empty()) {
            filteredStorage = readingIn;
            this->readingDataPointsFilters.init();
            this->readingDataPointsFilters.apply(filteredStorage);
        }
        const DataPoints& reading = this->readingDataPointsFilters.empty() ? readingIn : filteredStorage;
        struct KeepCopy {  // joins on every exit path, exceptions included
            std::future<void> f;
            ~KeepCopy() { if (f.valid()) f.get(); }
        } keep;
        keep.f = std::async(std::launch::async, [this, &reading] { readingFiltered = reading; });
        requireFloat3D(reading.features.rows(), "ICP");
        this->prefilteredReadingPtsCount = reading.features.cols();

        // T_refIn_refMean is a pure translation: its inverse negates it (ICP.cpp:345-347)
        const TransformationParameters T_refMean_dataIn = T_refMean_refIn * T_refIn_dataIn;
        RigidTransformation rigid;
        if (!rigid.checkParameters(T_refMean_dataIn)) throw TransformationError("RigidTransformation: Error, rotation matrix is not orthogonal.");

        auto* gpuMatcher = dynamic_cast<KDTreeMatcher*>(this->matcher.get());
        auto* gpuMinimizer = dynamic_cast<GpuErrorMinimizer*>(this->errorMinimizer.get());
        if (!gpuMatcher || !gpuMinimizer || !this->outlierFilters.allGpu() || this->outlierFilters.size() > 8)
            throw ConfigurationError("ICP: GPU build: matcher, outlier filters and error minimizer must all be GPU modules (there is no CPU path)");
        GpuPipeline& g = *pipeline;

        // the reading lives on the device, expressed in the refMean frame
        g.check(pmgpu_reading_set(g.ctx, reinterpret_cast<const float*>(reading.features.data()), reading.features.rows(), reading.features.cols()));
        gpuMatcher->uploadMaxDists(g, reading);  // KDTreeVarDistMatcher
        if (reading.descriptorExists("normals"))  // they turn with the reading (TransformationsImpl.cpp:71-84)
            g.check(pmgpu_reading_set_normals(g.ctx, reinterpret_cast<const float*>(reading.descriptors.data()) + reading.getDescriptorStartingRow("normals"),
                                              reading.descriptors.rows()));
        g.check(pmgpu_reading_apply_transform(g.ctx, reinterpret_cast<const float*>(T_refMean_dataIn.data())));
        g.readingKey = nullptr;

        TransformationParameters T_iter = Matrix::Identity(dim, dim);
        bool iterate = true;
        this->maxNumIterationsReached = false;
        iterationCount = 0;

        // which checkers are in the chain?
        CounterTransformationChecker* counter = nullptr;
        DifferentialTransformationChecker* differential = nullptr;
        bool onlyDeviceCheckers = true;
        for (auto& c : this->transformationCheckers) {
            if (auto* cc = dynamic_cast<CounterTransformationChecker*>(c.get())) counter = counter ? (onlyDeviceCheckers = false, counter) : cc;
            else if (auto* dc = dynamic_cast<DifferentialTransformationChecker*>(c.get())) differential = differential ? (onlyDeviceCheckers = false, differential) : dc;
            else onlyDeviceCheckers = false;
        }
        // chain order matters when Counter throws before Differential runs; the device evaluates
        // Counter first, so require that order (it is the order of every reference config)
        if (counter && differential && this->transformationCheckers[0].get() != counter) onlyDeviceCheckers = false;
        if (differential && differential->smoothLength >= 64) onlyDeviceCheckers = false;

        pmgpu_icp_params p;
        std::memset(&p, 0, sizeof(p));
        p.knn = gpuMatcher->knn;
        p.epsilon = (float)gpuMatcher->epsilon;
        p.max_dist = (float)gpuMatcher->maxDist;
        p.nfilters = this->outlierFilters.spec(p.filter_type, p.filter_param, g);
        p.minimizer = gpuMinimizer->kind;
        p.sensor_std_dev = (float)gpuMinimizer->sensorStdDev;
        p.max_iterations = counter ? (int)counter->maxIterationCount : 0x7fffffff;
        p.use_differential = differential ? 1 : 0;
        if (differential) {
            p.min_diff_rot_err = (float)differential->minDiffRotErr;
            p.min_diff_trans_err = (float)differential->minDiffTransErr;
            p.smooth_length = (int)differential->smoothLength;
        }

        fusedLoop = onlyDeviceCheckers && (counter || differential) && this->readingStepDataPointsFilters.empty() && !this->inspector->needsIterationData();
        float cov[36], stats[5];
        if (fusedLoop) {
            int iterations = 0;
            g.check(pmgpu_icp_run(g.ctx, &p, nullptr, reinterpret_cast<float*>(T_iter.data()), &iterations, cov, stats));
            iterationCount = iterations;
            this->maxNumIterationsReached = counter && iterations >= std::max(1, (int)counter->maxIterationCount);
            gpuMinimizer->setResults(cov, stats);
        } else {
            if (!this->readingStepDataPointsFilters.empty())
                throw ConfigurationError("ICP: GPU build: readingStepDataPointsFilters are not supported (the reading stays on the device)");
            this->transformationCheckers.init(T_iter, iterate);
            const bool wantHostData = this->inspector->needsIterationData();
            // what Inspector::dumpIteration is shown (ICP.cpp:345-347, 381, 403-405): the reading as the iteration sees it and the
            // reference in the frame of its mean — host copies, made only for an inspector that is not the NullInspector
            DataPoints readingRefMean, referenceCentred;
            if (wantHostData) {
                readingRefMean = RigidTransformation::apply(reading, T_refMean_dataIn);
                referenceCentred = reference;  // minus the mean, as ICP.cpp:291-299 subtracts it (a pure translation)
                for (int j = 0; j < referenceCentred.features.cols(); ++j)
                    for (int r = 0; r < dim - 1; ++r) referenceCentred.features(r, j) = referenceCentred.features(r, j) + T_refMean_refIn(r, dim - 1);
                if (!referenceCentred.descriptorExists("normals")) {  // made on the device by a trailing SurfaceNormal filter
                    Matrix normals(dim - 1, referenceCentred.features.cols());
                    if (pmgpu_ref_get_normals(g.ctx, reinterpret_cast<float*>(normals.data())) == PMGPU_OK) referenceCentred.addDescriptor("normals", normals);
                }
            }
            while (iterate) {
                // one iteration, stage by stage through the C ABI (ICP.cpp:371-430); host copies of
                // matches / weights are only made when an inspector wants to see them
                Matches matches(wantHostData ? p.knn : 0, wantHostData ? reading.features.cols() : 0);
                OutlierWeights weights(wantHostData ? p.knn : 0, wantHostData ? reading.features.cols() : 0);
                uint64_t visits = 0;
                g.check(pmgpu_knn(g.ctx, reinterpret_cast<const float*>(T_iter.data()), p.knn, p.epsilon, p.max_dist, wantHostData ? matches.ids.data() : nullptr,
                                  wantHostData ? reinterpret_cast<float*>(matches.dists.data()) : nullptr, &visits));
                gpuMatcher->visitCounter += visits;
                g.check(pmgpu_weights(g.ctx, p.nfilters, p.filter_type, p.filter_param, wantHostData ? reinterpret_cast<float*>(weights.data()) : nullptr, nullptr));
                if (wantHostData)
                    this->inspector->dumpIteration(iterationCount, T_iter, referenceCentred, RigidTransformation::apply(readingRefMean, T_iter), matches, weights,
                                                   this->transformationCheckers);
                TransformationParameters dT(dim, dim);
                g.check(pmgpu_minimize(g.ctx, p.minimizer, p.sensor_std_dev, reinterpret_cast<float*>(dT.data()), cov, stats));
                gpuMinimizer->setResults(cov, stats);
                T_iter = dT * T_iter;
                try {
                    this->transformationCheckers.check(T_iter, iterate);
                } catch (const typename CounterTransformationChecker::MaxNumIterationsReached&) {
                    iterate = false;
                    this->maxNumIterationsReached = true;
                }
                ++iterationCount;
                if (iterate && !rigid.checkParameters(T_iter)) throw TransformationError("RigidTransformation: Error, rotation matrix is not orthogonal.");
            }
        }
        this->inspector->addStat("IterationsCount", (double)iterationCount);
        this->inspector->addStat("PointCountTouched", (double)this->matcher->getVisitCount());
        this->matcher->resetVisitCount();
        this->inspector->addStat("OverlapRatio", this->errorMinimizer->getWeightedPointUsedRatio());
        this->inspector->finish(iterationCount);
        // ErrorElements of the last executed iteration, on request (ErrorMinimizer.cpp:58-193): the reading as that iteration saw
        // it (T_match * T_refMean_dataIn * filtered reading, ICP.cpp:345-347,381), the centred reference (device-made normals are
        // fetched), the kept pairs.  `reference` is the caller's cloud when there are no host reference filters: like the
        // reference's matcher (MatchersImpl.cpp:77-83, libnabo keeps a reference), it must stay alive while this is used.
        const DataPoints* referencePtr = &reference;
        const int knnUsed = p.knn;
        gpuMinimizer->materialize = [this, referencePtr, T_refMean_dataIn, T_refMean_refIn, knnUsed](typename ErrorMinimizer::ErrorElements& out) {
            GpuPipeline& gp = *pipeline;
            const int n = readingFiltered.features.cols();
            Matches m(knnUsed, n);
            OutlierWeights w(knnUsed, n);
            TransformationParameters T_match(readingFiltered.features.rows(), readingFiltered.features.rows());
            gp.check(pmgpu_matches_get(gp.ctx, m.ids.data(), reinterpret_cast<float*>(m.dists.data()), reinterpret_cast<float*>(w.data()),
                                       reinterpret_cast<float*>(T_match.data())));
            const DataPoints step = RigidTransformation::apply(RigidTransformation::apply(readingFiltered, T_refMean_dataIn), T_match);
            DataPoints centred = *referencePtr;
            for (int j = 0; j < centred.features.cols(); ++j)
                for (int r = 0; r < centred.features.rows() - 1; ++r) centred.features(r, j) = centred.features(r, j) + T_refMean_refIn(r, centred.features.rows() - 1);
            if (!centred.descriptorExists("normals")) {
                Matrix normals(centred.features.rows() - 1, centred.features.cols());
                if (pmgpu_ref_get_normals(gp.ctx, reinterpret_cast<float*>(normals.data())) == PMGPU_OK) centred.addDescriptor("normals", normals);
            }
            out = typename ErrorMinimizer::ErrorElements(step, centred, w, m);
        };
        // ICP.cpp:448
        return T_refIn_refMean * T_iter * T_refMean_dataIn;
    }
};

// ICPSequence (PointMatcher.h:726-764, ICP.cpp:455-609): ICP against a map that stays resident on
// the device.  setMap() centres, filters and indexes the map once; every operator() registers a
// new cloud against it (BASELINE config 3: a 10 M-point map, readings streaming in).
struct ICPSequence : public ICP {
    bool hasMap() const { return mapPointCloud.features.cols() != 0; }
    // ICP.cpp:463-508
    bool setMap(const DataPoints& inputCloud) {
        if (!this->matcher) throw std::runtime_error("You must setup a matcher before running ICP");
        if (!this->inspector) throw std::runtime_error("You must setup an inspector before running ICP");
        const int dim = inputCloud.features.rows();
        if (inputCloud.features.cols() == 0) return false;
        requireFloat3D(dim, "ICPSequence");
        this->bindPipeline();
        auto* gpuMatcher = dynamic_cast<KDTreeMatcher*>(this->matcher.get());
        if (!gpuMatcher) throw ConfigurationError("ICPSequence: GPU build: the matcher must be the GPU KDTreeMatcher (there is no CPU path)");
        this->inspector->addStat("MapPointCount", inputCloud.features.cols());
        mapPointCloud = inputCloud;
        // the reference centres first and filters afterwards (ICP.cpp:486-496); the GPU filters on
        // this path (SurfaceNormal, Identity) do not move points, so filtering the un-centred
        // cloud and centring inside the matcher init is the same computation
        this->referenceDataPointsFilters.init();
        this->referenceDataPointsFilters.apply(mapPointCloud);
        float mean4[4];
        gpuMatcher->initCentered(mapPointCloud, mean4);
        T_refIn_refMean = Matrix::Identity(dim, dim);
        T_refMean_refIn = Matrix::Identity(dim, dim);
        for (int r = 0; r < dim - 1; ++r) {
            T_refIn_refMean(r, dim - 1) = T(mean4[r]);
            T_refMean_refIn(r, dim - 1) = T(-mean4[r]);
        }
        this->prefilteredReferencePtsCount = mapPointCloud.features.cols();
        return true;
    }
    void clearMap() { mapPointCloud = DataPoints(); }
    // the map in its original frame (the host copy is kept un-centred here)
    const DataPoints& getPrefilteredMap() const { return mapPointCloud; }
    const DataPoints& getMap() const { return mapPointCloud; }

    TransformationParameters operator()(const DataPoints& cloudIn) {
        const int dim = cloudIn.features.rows();
        return this->compute(cloudIn, Matrix::Identity(dim, dim));
    }
    TransformationParameters operator()(const DataPoints& cloudIn, const TransformationParameters& T_dataInOld_dataInNew) {
        return this->compute(cloudIn, T_dataInOld_dataInNew);
    }
    // ICP.cpp:595-609
    TransformationParameters compute(const DataPoints& cloudIn, const TransformationParameters& T_refIn_dataIn) {
        if (!hasMap()) {
            const int dim = cloudIn.features.rows();
            return Matrix::Identity(dim, dim);  // "Ignoring attempt to perform ICP with an empty map"
        }
        this->bindPipeline();
        this->inspector->init();
        return this->computeWithTransformedReference(cloudIn, mapPointCloud, T_refIn_refMean, T_refMean_refIn, T_refIn_dataIn);
    }

protected:
    DataPoints mapPointCloud;
    TransformationParameters T_refIn_refMean, T_refMean_refIn;
};

